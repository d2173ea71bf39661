"""cmdstan_b200 -- B200-native NUTS (diag_e, adaptive) engine behind CmdStan's sampling surface.

Python host side of the C ABI in include/b200nuts.h (ctypes; no torch types cross the boundary).  The public call mirrors
the reference service stan::services::sample::hmc_nuts_diag_e_adapt (stan/src/stan/services/sample/
hmc_nuts_diag_e_adapt.hpp:331-407): same argument names, meaning, defaults and error codes.

There is no CPU path: importing works anywhere (the shared library loads without a driver), but every compute call
fails loudly with error code 70 when no CUDA device is present.
"""
import ctypes as C
import os

import numpy as np

from . import _build

__all__ = ["Engine", "hmc_nuts_diag_e_adapt", "hmc_nuts_diag_e", "hmc_nuts_unit_e_adapt", "hmc_nuts_unit_e", "hmc_nuts_dense_e_adapt", "hmc_nuts_dense_e", "hmc_static_diag_e_adapt", "hmc_static_diag_e", "hmc_static_unit_e_adapt",
           "hmc_static_unit_e", "hmc_static_dense_e_adapt", "hmc_static_dense_e", "STATIC_SAMPLER_COLUMNS", "param_names", "B200NutsError", "MODELS", "load_library", "SAMPLER_COLUMNS", "SUMMARY_COLUMNS", "summarize_draws", "nccl_unique_id", "load_model_library", "find_model"]

MODELS = {"std_normal": 0, "bernoulli": 1, "eight_schools_nc": 2, "logistic": 3, "irt_2pl": 4, "poisson_glm": 5}
ENGINES = {"auto": 0, "fused": 1, "lockstep": 2}
SUMMARY_COLUMNS = ["Mean", "MCSE", "StdDev", "MAD", "5%", "50%", "95%", "ESS_bulk", "ESS_tail", "R_hat"]
SAMPLER_COLUMNS = ["lp__", "accept_stat__", "stepsize__", "treedepth__", "n_leapfrog__", "divergent__", "energy__"]
STATIC_SAMPLER_COLUMNS = ["lp__", "accept_stat__", "stepsize__", "int_time__", "energy__"]   # base_static_hmc::get_sampler_param_names

# stan::services::error_codes (stan/src/stan/services/error_codes.hpp)
OK, USAGE, DATAERR, NOINPUT, SOFTWARE, CONFIG = 0, 64, 65, 66, 70, 78


class Config(C.Structure):
    _fields_ = [("model", C.c_int32), ("num_chains", C.c_int32), ("random_seed", C.c_uint32), ("init_chain_id", C.c_uint32),
                ("init_radius", C.c_double), ("num_warmup", C.c_int32), ("num_samples", C.c_int32), ("num_thin", C.c_int32),
                ("save_warmup", C.c_int32), ("stepsize", C.c_double), ("stepsize_jitter", C.c_double), ("max_depth", C.c_int32),
                ("delta", C.c_double), ("gamma", C.c_double), ("kappa", C.c_double), ("t0", C.c_double),
                ("init_buffer", C.c_uint32), ("term_buffer", C.c_uint32), ("window", C.c_uint32), ("adapt_engaged", C.c_int32),
                ("engine", C.c_int32), ("device", C.c_int32), ("store_draws", C.c_int32), ("row_shards", C.c_int32),
                ("row_shard_rank", C.c_int32), ("metric", C.c_int32), ("store_diagnostics", C.c_int32), ("static_hmc", C.c_int32), ("int_time", C.c_double)]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p)

EXPORTS = ["b200nuts_default_config", "b200nuts_create", "b200nuts_destroy", "b200nuts_last_error", "b200nuts_set_data_f64",
           "b200nuts_set_data_i32", "b200nuts_set_data_i8", "b200nuts_num_params", "b200nuts_set_inits", "b200nuts_set_inv_metric",
           "b200nuts_init_chains", "b200nuts_log_prob_grad", "b200nuts_run", "b200nuts_fetch_draws", "b200nuts_fetch_diagnostics", "b200nuts_rows_saved",
           "b200nuts_get_adaptation", "b200nuts_get_chain_status", "b200nuts_get_timing", "b200nuts_set_allreduce",
           "b200nuts_nccl_get_unique_id", "b200nuts_nccl_init", "b200nuts_measure_allreduce", "b200nuts_measure_fp64_fma_peak", "b200nuts_summary", "b200nuts_summary_num_stats", "b200nuts_summarize_draws",
           # include/b200nuts_model.h
           "b200nuts_register_model", "b200nuts_load_model_library", "b200nuts_find_model", "b200nuts_output_name", "b200nuts_transform_inits"]

_lib = None


class B200NutsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"b200nuts error {code}: {msg}")
        self.code = code


def load_library(build=True):
    """Loads (building first if sources changed) cmdstan_b200/libb200nuts.so.  Raises if it is missing: no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if build and _build.needs_build():
        _build.build()
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: run `python -m cmdstan_b200._build` (nvcc, sm_100a); there is no CPU fallback")
    lib = C.CDLL(path)
    vp, i64p = C.c_void_p, C.POINTER(C.c_int64)
    lib.b200nuts_default_config.argtypes = [C.POINTER(Config)]
    lib.b200nuts_default_config.restype = None
    lib.b200nuts_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    lib.b200nuts_destroy.argtypes = [vp]
    lib.b200nuts_destroy.restype = None
    lib.b200nuts_last_error.argtypes = [vp]
    lib.b200nuts_last_error.restype = C.c_char_p
    for f in ("b200nuts_set_data_f64", "b200nuts_set_data_i32", "b200nuts_set_data_i8"):
        getattr(lib, f).argtypes = [vp, C.c_char_p, vp, i64p, C.c_int]
    lib.b200nuts_num_params.argtypes = [vp, i64p, i64p]
    lib.b200nuts_set_inits.argtypes = [vp, vp]
    lib.b200nuts_set_inv_metric.argtypes = [vp, vp]
    lib.b200nuts_init_chains.argtypes = [vp]
    lib.b200nuts_log_prob_grad.argtypes = [vp, vp, C.c_int64, C.c_int, vp, vp]
    lib.b200nuts_run.argtypes = [vp, C.c_int]
    lib.b200nuts_fetch_draws.argtypes = [vp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int, vp]
    lib.b200nuts_fetch_diagnostics.argtypes = [vp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, vp]
    lib.b200nuts_rows_saved.argtypes = [vp]
    lib.b200nuts_rows_saved.restype = C.c_int64
    lib.b200nuts_get_adaptation.argtypes = [vp, vp, vp]
    lib.b200nuts_get_chain_status.argtypes = [vp, vp, vp, vp]
    lib.b200nuts_get_timing.argtypes = [vp, C.POINTER(C.c_double), i64p, C.POINTER(C.c_double)]
    lib.b200nuts_set_allreduce.argtypes = [vp, ALLREDUCE_FN, vp]
    lib.b200nuts_nccl_get_unique_id.argtypes = [vp]
    lib.b200nuts_nccl_init.argtypes = [vp, vp, C.c_int, C.c_int]
    lib.b200nuts_measure_allreduce.argtypes = [vp, C.c_int, C.POINTER(C.c_double), i64p]
    lib.b200nuts_summary.argtypes = [vp, C.c_int64, C.c_int64, C.c_int, vp]
    lib.b200nuts_summarize_draws.argtypes = [vp, C.c_int64, C.c_int64, C.c_int64, C.c_int, vp, C.c_char_p, C.c_int]
    lib.b200nuts_measure_fp64_fma_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
    lib.b200nuts_register_model.argtypes = [vp]
    lib.b200nuts_load_model_library.argtypes = [C.c_char_p]
    lib.b200nuts_find_model.argtypes = [C.c_char_p]
    lib.b200nuts_output_name.argtypes = [vp, C.c_int64, C.c_char_p, C.c_int]
    lib.b200nuts_transform_inits.argtypes = [vp, vp, C.c_int64, vp]
    _lib = lib
    return lib


def load_model_library(path):
    """Registers the device models of a plug-in library (include/b200nuts_model.h: a .so built with nvcc that exports
    b200nuts_model_entry) and returns the id of its first model.  The device-side counterpart of linking a stanc-generated
    model into CmdStan (new_model, src/cmdstan/command.hpp:71-73)."""
    lib = load_library()
    rc = lib.b200nuts_load_model_library(os.fsencode(path))
    if rc < 0:
        raise B200NutsError(-rc, lib.b200nuts_last_error(None).decode())
    return rc


def measure_fp64_fma_peak(device=0):
    """Measured FP64 FMA rate of the device in TFLOP/s (roofline denominator of the scalar-FP64-bound kernels)."""
    lib = load_library()
    v = C.c_double()
    rc = lib.b200nuts_measure_fp64_fma_peak(int(device), C.byref(v))
    if rc:
        raise B200NutsError(rc, lib.b200nuts_last_error(None).decode())
    return v.value


def find_model(name):
    """id of a built-in or registered plug-in model by model_name(); raises KeyError when unknown."""
    rc = load_library().b200nuts_find_model(name.encode())
    if rc < 0:
        raise KeyError(name)
    return rc


def _point_at_bundled_nccl():
    """The torch wheel bundles NCCL (site-packages/nvidia/nccl/lib/libnccl.so.2); tell the library where it is."""
    if "B200NUTS_NCCL_LIB" in os.environ:
        return
    try:
        import nvidia.nccl
        for d in nvidia.nccl.__path__:
            p = os.path.join(d, "lib", "libnccl.so.2")
            if os.path.exists(p):
                os.environ["B200NUTS_NCCL_LIB"] = p
                return
    except Exception:
        pass


def nccl_unique_id():
    """128-byte NCCL unique id (rank 0 creates it and broadcasts it, e.g. with torch.distributed.broadcast_object_list)."""
    _point_at_bundled_nccl()
    lib = load_library()
    buf = C.create_string_buffer(128)
    rc = lib.b200nuts_nccl_get_unique_id(buf)
    if rc:
        raise B200NutsError(rc, lib.b200nuts_last_error(None).decode())
    return buf.raw


def summarize_draws(draws, device=0):
    """stansummary's table (columns SUMMARY_COLUMNS) of host draws [chains, rows, ncols], computed on the GPU."""
    lib = load_library()
    a = np.ascontiguousarray(draws, dtype=np.float64)
    assert a.ndim == 3
    out = np.empty((a.shape[2], lib.b200nuts_summary_num_stats()))
    err = C.create_string_buffer(512)
    rc = lib.b200nuts_summarize_draws(a.ctypes.data_as(C.c_void_p), a.shape[0], a.shape[1], a.shape[2], int(device),
                                      out.ctypes.data_as(C.c_void_p), err, 512)
    if rc:
        raise B200NutsError(rc, err.value.decode())
    return out


def _np(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


class Engine:
    """One engine = one GPU (one rank): owns the model data and the state of `num_chains` chains on `device`."""

    def __init__(self, model, data, num_chains=1, random_seed=0, init_chain_id=1, init_radius=2.0, num_warmup=1000,
                 num_samples=1000, num_thin=1, save_warmup=False, stepsize=1.0, stepsize_jitter=0.0, max_depth=10, delta=0.8,
                 gamma=0.05, kappa=0.75, t0=10.0, init_buffer=75, term_buffer=50, window=25, adapt_engaged=True, engine="auto",
                 device=0, store_draws=True, row_shards=1, row_shard_rank=0, allreduce=None, metric="diag_e", store_diagnostics=False, static_hmc=False, int_time=2 * np.pi):
        self.lib = load_library()
        cfg = Config()
        self.lib.b200nuts_default_config(C.byref(cfg))
        cfg.model = (MODELS[model] if model in MODELS else find_model(model)) if isinstance(model, str) else int(model)
        cfg.num_chains, cfg.random_seed, cfg.init_chain_id = int(num_chains), int(random_seed), int(init_chain_id)
        cfg.init_radius, cfg.num_warmup, cfg.num_samples, cfg.num_thin = float(init_radius), int(num_warmup), int(num_samples), int(num_thin)
        cfg.save_warmup, cfg.stepsize, cfg.stepsize_jitter, cfg.max_depth = int(bool(save_warmup)), float(stepsize), float(stepsize_jitter), int(max_depth)
        cfg.delta, cfg.gamma, cfg.kappa, cfg.t0 = float(delta), float(gamma), float(kappa), float(t0)
        cfg.init_buffer, cfg.term_buffer, cfg.window = int(init_buffer), int(term_buffer), int(window)
        cfg.adapt_engaged, cfg.engine, cfg.device, cfg.store_draws = int(bool(adapt_engaged)), ENGINES[engine], int(device), int(bool(store_draws))
        cfg.row_shards, cfg.row_shard_rank = int(row_shards), int(row_shard_rank)
        cfg.metric = {"diag_e": 0, "unit_e": 1, "dense_e": 2}[metric]
        cfg.store_diagnostics = int(bool(store_diagnostics))
        cfg.static_hmc, cfg.int_time = int(bool(static_hmc)), float(int_time)
        self.metric = metric
        self.cfg = cfg
        self.model = model
        self.h = C.c_void_p()
        rc = self.lib.b200nuts_create(C.byref(cfg), C.byref(self.h))
        if rc:
            raise B200NutsError(rc, self.lib.b200nuts_last_error(None).decode())
        self._cb = None
        try:
            self._set_data(data or {})
            nu, no = C.c_int64(), C.c_int64()
            self._chk(self.lib.b200nuts_num_params(self.h, C.byref(nu), C.byref(no)))
            self.D, self.K = nu.value, no.value
            if allreduce is not None:
                self._cb = ALLREDUCE_FN(allreduce)
                self._chk(self.lib.b200nuts_set_allreduce(self.h, self._cb, None))
        except Exception:
            self.close()
            raise
        self.num_chains = int(num_chains)

    # -- plumbing
    def _chk(self, rc):
        if rc:
            raise B200NutsError(rc, self.lib.b200nuts_last_error(self.h).decode())

    def _set_data(self, data):
        for name, val in data.items():
            a = np.asarray(val)
            if a.dtype == np.int8:
                fn, a = self.lib.b200nuts_set_data_i8, _np(a, np.int8)
            elif np.issubdtype(a.dtype, np.integer) or a.dtype == np.bool_:
                fn, a = self.lib.b200nuts_set_data_i32, _np(a, np.int32)
            else:
                fn, a = self.lib.b200nuts_set_data_f64, _np(a, np.float64)
            if self.model in ("eight_schools_nc", 2) and name in ("y", "sigma"):
                fn, a = self.lib.b200nuts_set_data_f64, _np(a, np.float64)
            if self.model in ("irt_2pl", 4) and name == "y":
                fn, a = self.lib.b200nuts_set_data_i8, _np(a, np.int8)
            dims = (C.c_int64 * max(a.ndim, 1))(*(a.shape if a.ndim else (1,)))
            self._chk(fn(self.h, name.encode(), a.ctypes.data_as(C.c_void_p), dims, max(a.ndim, 1)))

    def nccl_init(self, unique_id, nranks, rank):
        """Row sharding: join the NCCL communicator identified by `unique_id` (128 bytes from nccl_unique_id() on rank 0);
        from then on every gradient stage all-reduces its (lp, gradient) partials with ncclAllReduce on the engine stream."""
        _point_at_bundled_nccl()
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._chk(self.lib.b200nuts_nccl_init(self.h, buf, int(nranks), int(rank)))

    def measure_allreduce(self, iters=200):
        """row-sharded engines: (average microseconds, payload bytes) of the per-leapfrog all-reduce, `iters` of them back to
        back on the engine's stream (collective: call on every rank)."""
        us, nb = C.c_double(), C.c_int64()
        self._chk(self.lib.b200nuts_measure_allreduce(self.h, int(iters), C.byref(us), C.byref(nb)))
        return us.value, nb.value

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.b200nuts_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- the path
    def set_inits(self, q0):
        a = None if q0 is None else _np(q0, np.float64).reshape(self.num_chains, self.D)
        self._chk(self.lib.b200nuts_set_inits(self.h, None if a is None else a.ctypes.data_as(C.c_void_p)))

    def transform_inits(self, x):
        """model_base::transform_inits: initial values on the constrained scale, x [n, K] in write_array's column order (NaN = not
        given) -> unconstrained q [n, D] for set_inits (built-in models and plug-in models with vtable.transform_inits)."""
        x = _np(x, np.float64).reshape(-1, self.K)
        q = np.empty((x.shape[0], self.D))
        self._chk(self.lib.b200nuts_transform_inits(self.h, x.ctypes.data_as(C.c_void_p), x.shape[0], q.ctypes.data_as(C.c_void_p)))
        return q

    def set_inv_metric(self, im):
        shape = (self.num_chains, self.D, self.D) if self.metric == "dense_e" else (self.num_chains, self.D)
        a = None if im is None else _np(im, np.float64).reshape(shape)
        self._chk(self.lib.b200nuts_set_inv_metric(self.h, None if a is None else a.ctypes.data_as(C.c_void_p)))

    def init_chains(self):
        self._chk(self.lib.b200nuts_init_chains(self.h))

    def log_prob_grad(self, q, jacobian=True):
        q = _np(q, np.float64)
        q2 = q.reshape(-1, self.D)
        lp = np.empty(len(q2))
        g = np.empty_like(q2)
        self._chk(self.lib.b200nuts_log_prob_grad(self.h, q2.ctypes.data_as(C.c_void_p), len(q2), int(jacobian),
                                                  lp.ctypes.data_as(C.c_void_p), g.ctypes.data_as(C.c_void_p)))
        return (lp[0], g[0]) if q.ndim == 1 else (lp, g)

    def run(self, n_iter):
        self._chk(self.lib.b200nuts_run(self.h, int(n_iter)))

    def rows_saved(self):
        return int(self.lib.b200nuts_rows_saved(self.h))

    def fetch_draws(self, chain0=0, nchains=None, row0=0, nrows=None, constrained=True, out=None):
        nchains = self.num_chains - chain0 if nchains is None else nchains
        nrows = self.rows_saved() - row0 if nrows is None else nrows
        k = self.K if constrained else self.D
        if out is None:
            out = np.empty((nchains, nrows, 7 + k))
        self._chk(self.lib.b200nuts_fetch_draws(self.h, chain0, nchains, row0, nrows, int(constrained), out.ctypes.data_as(C.c_void_p)))
        return out

    def fetch_diagnostics(self, chain0=0, nchains=None, row0=0, nrows=None):
        """[nchains, nrows, 7 + 3 D]: sampler columns, unconstrained q, momentum p, gradient of the potential g of every saved
        draw (the rows of CmdStan's `output diagnostic_file=`); needs store_diagnostics=True."""
        nchains = self.num_chains - chain0 if nchains is None else nchains
        nrows = self.rows_saved() - row0 if nrows is None else nrows
        out = np.empty((nchains, nrows, 7 + 3 * self.D))
        self._chk(self.lib.b200nuts_fetch_diagnostics(self.h, chain0, nchains, row0, nrows, out.ctypes.data_as(C.c_void_p)))
        return out

    def summary(self, row0=0, nrows=None, constrained=True):
        """stansummary's table computed on the device from the stored draws: [7 + K, 10] with columns SUMMARY_COLUMNS."""
        nrows = self.rows_saved() - row0 if nrows is None else nrows
        k = self.K if constrained else self.D
        out = np.empty((7 + k, self.lib.b200nuts_summary_num_stats()))
        self._chk(self.lib.b200nuts_summary(self.h, row0, nrows, int(constrained), out.ctypes.data_as(C.c_void_p)))
        return out

    def output_names(self):
        """constrained_param_names() of the engine's model (built-in or plug-in)."""
        buf = C.create_string_buffer(256)
        names = []
        for k in range(self.K):
            self._chk(self.lib.b200nuts_output_name(self.h, k, buf, 256))
            names.append(buf.value.decode())
        return names

    def adaptation(self):
        eps = np.empty(self.num_chains)
        im = np.empty((self.num_chains, self.D, self.D) if self.metric == "dense_e" else (self.num_chains, self.D))
        self._chk(self.lib.b200nuts_get_adaptation(self.h, eps.ctypes.data_as(C.c_void_p), im.ctypes.data_as(C.c_void_p)))
        return eps, im

    def chain_status(self):
        err = np.empty(self.num_chains, dtype=np.int32)
        it = np.empty(self.num_chains, dtype=np.int32)
        ng = np.empty(self.num_chains, dtype=np.int64)
        self._chk(self.lib.b200nuts_get_chain_status(self.h, err.ctypes.data_as(C.c_void_p), it.ctypes.data_as(C.c_void_p), ng.ctypes.data_as(C.c_void_p)))
        return err, it, ng

    def timing(self):
        s, n, g = C.c_double(), C.c_int64(), C.c_double()
        self._chk(self.lib.b200nuts_get_timing(self.h, C.byref(s), C.byref(n), C.byref(g)))
        return s.value, n.value, g.value


def param_names(model, data):
    """constrained_param_names(include_tparams=True) of the device model family (host mirror: cmdstan_b200/host/model.hpp)."""
    vec = lambda n, k: [f"{n}.{i}" for i in range(1, k + 1)]
    if model == "std_normal":
        return vec("x", int(np.asarray(data.get("D", 3)).ravel()[0]))
    if model == "bernoulli":
        return ["theta"]
    if model == "eight_schools_nc":
        J = len(data["y"])
        return ["mu", "tau"] + vec("theta_tilde", J) + vec("theta", J)
    if model in ("logistic", "poisson_glm"):
        return ["alpha"] + vec("beta", np.asarray(data["X"]).shape[1])
    if model == "irt_2pl":
        I, J = np.asarray(data["y"]).shape
        return ["sigma_theta"] + vec("theta", I) + ["sigma_a"] + vec("a", J) + ["mu_b", "sigma_b"] + vec("b", J)
    raise KeyError(model)


def _run_service(model, data, num_chains, init, init_inv_metric, random_seed, init_chain_id, init_radius, num_warmup, num_samples,
                 num_thin, save_warmup, stepsize, stepsize_jitter, max_depth, delta, gamma, kappa, t0, init_buffer, term_buffer,
                 window, logger, device, engine, adapt_engaged, metric, int_time=None):
    log = logger or (lambda msg: None)
    try:
        e = Engine(model, data, num_chains=num_chains, random_seed=random_seed, init_chain_id=init_chain_id,
                   init_radius=init_radius, num_warmup=num_warmup, num_samples=num_samples, num_thin=num_thin,
                   save_warmup=save_warmup, stepsize=stepsize, stepsize_jitter=stepsize_jitter, max_depth=max_depth, delta=delta,
                   gamma=gamma, kappa=kappa, t0=t0, init_buffer=init_buffer, term_buffer=term_buffer, window=window,
                   adapt_engaged=adapt_engaged, engine=engine, device=device, metric=metric, static_hmc=int_time is not None,
                   int_time=int_time if int_time is not None else 2 * np.pi)
    except B200NutsError as ex:
        log(str(ex))
        return ex.code, None
    try:
        if init is not None:
            e.set_inits(init)
        if init_inv_metric is not None:
            e.set_inv_metric(init_inv_metric)
        e.init_chains()
        e.run(num_warmup)
        warm_s = e.timing()[0]
        e.run(num_samples)
        samp_s = e.timing()[0]
        err, it, ng = e.chain_status()
        if (err != 0).any():
            code = int(err[err != 0][0])
            log("chain(s) %s stopped: %s" % (np.nonzero(err)[0].tolist(), {78: "Initialization failed.", 70: "step size / metric adaptation failure"}.get(code, code)))
            return code, None
        eps, im = e.adaptation()
        draws = e.fetch_draws()
        names = SAMPLER_COLUMNS + e.output_names()
        if int_time is not None:   # static HMC: the engine's rows carry L, 1, 0 where NUTS has treedepth, n_leapfrog, divergent
            draws = np.concatenate([draws[:, :, :3], np.full(draws.shape[:2] + (1,), float(int_time)), draws[:, :, 6:]], axis=2)
            names = STATIC_SAMPLER_COLUMNS + e.output_names()
        return OK, dict(draws=draws, columns=names, stepsize=eps, inv_metric=im, timing=dict(warmup=warm_s, sampling=samp_s), n_grad=ng)
    except B200NutsError as ex:
        log(str(ex))
        return ex.code, None
    finally:
        e.close()


def hmc_nuts_diag_e_adapt(model, data, num_chains=1, init=None, init_inv_metric=None, random_seed=0, init_chain_id=1,
                          init_radius=2.0, num_warmup=1000, num_samples=1000, num_thin=1, save_warmup=False, refresh=100,
                          stepsize=1.0, stepsize_jitter=0.0, max_depth=10, delta=0.8, gamma=0.05, kappa=0.75, t0=10.0,
                          init_buffer=75, term_buffer=50, window=25, logger=None, device=0, engine="auto"):
    """Mirror of stan::services::sample::hmc_nuts_diag_e_adapt (multi-chain overload,
    stan/src/stan/services/sample/hmc_nuts_diag_e_adapt.hpp:331-407).

    Returns (return_code, result) with return_code in stan::services::error_codes; result is a dict with
    `draws` [num_chains, rows, 7 + K], `columns`, `stepsize` [num_chains], `inv_metric` [num_chains, D], `timing`.
    `init` are unconstrained initial values [num_chains, D] (None: uniform(-init_radius, init_radius) from each chain's
    own stream, like util::initialize); `init_inv_metric` [num_chains, D] (None: ones).
    """
    return _run_service(model, data, num_chains, init, init_inv_metric, random_seed, init_chain_id, init_radius, num_warmup, num_samples,
                        num_thin, save_warmup, stepsize, stepsize_jitter, max_depth, delta, gamma, kappa, t0, init_buffer, term_buffer,
                        window, logger, device, engine, True, "diag_e")


def hmc_nuts_diag_e(model, data, num_chains=1, init=None, init_inv_metric=None, random_seed=0, init_chain_id=1, init_radius=2.0,
                    num_warmup=1000, num_samples=1000, num_thin=1, save_warmup=False, refresh=100, stepsize=1.0,
                    stepsize_jitter=0.0, max_depth=10, logger=None, device=0, engine="auto"):
    """stan::services::sample::hmc_nuts_diag_e (stan/src/stan/services/sample/hmc_nuts_diag_e.hpp): fixed step size and
    metric, no adaptation (util::run_sampler: warm-up transitions are ordinary transitions, no init_stepsize)."""
    return _run_service(model, data, num_chains, init, init_inv_metric, random_seed, init_chain_id, init_radius, num_warmup, num_samples,
                        num_thin, save_warmup, stepsize, stepsize_jitter, max_depth, 0.8, 0.05, 0.75, 10.0, 75, 50, 25, logger, device,
                        engine, False, "diag_e")


def hmc_nuts_unit_e_adapt(model, data, num_chains=1, init=None, random_seed=0, init_chain_id=1, init_radius=2.0, num_warmup=1000,
                          num_samples=1000, num_thin=1, save_warmup=False, refresh=100, stepsize=1.0, stepsize_jitter=0.0,
                          max_depth=10, delta=0.8, gamma=0.05, kappa=0.75, t0=10.0, logger=None, device=0, engine="auto"):
    """stan::services::sample::hmc_nuts_unit_e_adapt (stan/src/stan/services/sample/hmc_nuts_unit_e_adapt.hpp): identity
    metric, step-size adaptation only (adapt_unit_e_nuts.hpp:24-35)."""
    return _run_service(model, data, num_chains, init, None, random_seed, init_chain_id, init_radius, num_warmup, num_samples, num_thin,
                        save_warmup, stepsize, stepsize_jitter, max_depth, delta, gamma, kappa, t0, 75, 50, 25, logger, device, engine,
                        True, "unit_e")


def hmc_nuts_unit_e(model, data, num_chains=1, init=None, random_seed=0, init_chain_id=1, init_radius=2.0, num_warmup=1000,
                    num_samples=1000, num_thin=1, save_warmup=False, refresh=100, stepsize=1.0, stepsize_jitter=0.0, max_depth=10,
                    logger=None, device=0, engine="auto"):
    """stan::services::sample::hmc_nuts_unit_e (stan/src/stan/services/sample/hmc_nuts_unit_e.hpp): identity metric, no adaptation."""
    return _run_service(model, data, num_chains, init, None, random_seed, init_chain_id, init_radius, num_warmup, num_samples, num_thin,
                        save_warmup, stepsize, stepsize_jitter, max_depth, 0.8, 0.05, 0.75, 10.0, 75, 50, 25, logger, device, engine,
                        False, "unit_e")


def hmc_nuts_dense_e_adapt(model, data, num_chains=1, init=None, init_inv_metric=None, random_seed=0, init_chain_id=1,
                           init_radius=2.0, num_warmup=1000, num_samples=1000, num_thin=1, save_warmup=False, refresh=100,
                           stepsize=1.0, stepsize_jitter=0.0, max_depth=10, delta=0.8, gamma=0.05, kappa=0.75, t0=10.0,
                           init_buffer=75, term_buffer=50, window=25, logger=None, device=0, engine="auto"):
    """stan::services::sample::hmc_nuts_dense_e_adapt (stan/src/stan/services/sample/hmc_nuts_dense_e_adapt.hpp): dense
    Euclidean metric, step size and covariance adapted in windows (adapt_dense_e_nuts.hpp:25-44).  `init_inv_metric` is
    [num_chains, D, D] (None: identity); the result's `inv_metric` has the same shape."""
    return _run_service(model, data, num_chains, init, init_inv_metric, random_seed, init_chain_id, init_radius, num_warmup, num_samples,
                        num_thin, save_warmup, stepsize, stepsize_jitter, max_depth, delta, gamma, kappa, t0, init_buffer, term_buffer,
                        window, logger, device, engine, True, "dense_e")


def hmc_nuts_dense_e(model, data, num_chains=1, init=None, init_inv_metric=None, random_seed=0, init_chain_id=1, init_radius=2.0,
                     num_warmup=1000, num_samples=1000, num_thin=1, save_warmup=False, refresh=100, stepsize=1.0,
                     stepsize_jitter=0.0, max_depth=10, logger=None, device=0, engine="auto"):
    """stan::services::sample::hmc_nuts_dense_e (stan/src/stan/services/sample/hmc_nuts_dense_e.hpp): fixed step size and
    dense metric, no adaptation."""
    return _run_service(model, data, num_chains, init, init_inv_metric, random_seed, init_chain_id, init_radius, num_warmup, num_samples,
                        num_thin, save_warmup, stepsize, stepsize_jitter, max_depth, 0.8, 0.05, 0.75, 10.0, 75, 50, 25, logger, device,
                        engine, False, "dense_e")


def _static_service(metric, adapt):
    def service(model, data, num_chains=1, init=None, init_inv_metric=None, random_seed=0, init_chain_id=1, init_radius=2.0, num_warmup=1000,
                num_samples=1000, num_thin=1, save_warmup=False, refresh=100, stepsize=1.0, stepsize_jitter=0.0, int_time=2 * np.pi, delta=0.8,
                gamma=0.05, kappa=0.75, t0=10.0, init_buffer=75, term_buffer=50, window=25, logger=None, device=0, engine="auto"):
        return _run_service(model, data, num_chains, init, None if metric == "unit_e" else init_inv_metric, random_seed, init_chain_id, init_radius,
                            num_warmup, num_samples, num_thin, save_warmup, stepsize, stepsize_jitter, 10, delta, gamma, kappa, t0, init_buffer,
                            term_buffer, window, logger, device, engine, adapt, metric, int_time=float(int_time))
    service.__name__ = "hmc_static_%s%s" % (metric, "_adapt" if adapt else "")
    service.__doc__ = ("stan::services::sample::%s (stan/src/stan/services/sample/%s.hpp): static HMC with integration time `int_time` "
                       "(stan/mcmc/hmc/static/base_static_hmc.hpp:41-70), %s metric, %s.  Result columns: STATIC_SAMPLER_COLUMNS + the model's."
                       % (service.__name__, service.__name__, metric, "step size (and metric) adapted in windows" if adapt else "no adaptation"))
    return service


hmc_static_diag_e_adapt, hmc_static_diag_e = _static_service("diag_e", True), _static_service("diag_e", False)
hmc_static_unit_e_adapt, hmc_static_unit_e = _static_service("unit_e", True), _static_service("unit_e", False)
hmc_static_dense_e_adapt, hmc_static_dense_e = _static_service("dense_e", True), _static_service("dense_e", False)
