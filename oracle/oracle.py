"""TEST INFRASTRUCTURE ONLY -- ctypes front end for the CPU oracle (liboracle.so) and the compiled reference (oracle/_ref).

May be imported from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs only.
The product package (cmdstan_b200) never imports this module.
"""
import ctypes as C
import json
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")

M_STD_NORMAL, M_BERNOULLI, M_EIGHT_SCHOOLS, M_LOGISTIC, M_IRT2PL, M_POISSON = range(6)
KIND_BY_NAME = {"std_normal": 0, "bernoulli": 1, "eight_schools_nc": 2, "logistic": 3, "irt_2pl": 4, "poisson_glm": 5,
                "mock_flat": 6}   # mock_flat: log_prob == 0, the mock_model of the reference's sampler unit tests (KATs only)


class ModelT(C.Structure):
    _fields_ = [("kind", C.c_int), ("D", C.c_int), ("N", C.c_int), ("K", C.c_int), ("X", C.c_void_p), ("yi", C.c_void_p),
                ("yr", C.c_void_p), ("sigma", C.c_void_p), ("y8", C.c_void_p)]


class ConfigT(C.Structure):
    _fields_ = [("seed", C.c_uint32), ("chain_id", C.c_uint32), ("num_warmup", C.c_int), ("num_samples", C.c_int),
                ("num_thin", C.c_int), ("save_warmup", C.c_int), ("stepsize", C.c_double), ("stepsize_jitter", C.c_double),
                ("max_depth", C.c_int), ("delta", C.c_double), ("gamma", C.c_double), ("kappa", C.c_double),
                ("t0", C.c_double), ("init_buffer", C.c_uint), ("term_buffer", C.c_uint), ("window", C.c_uint),
                ("init_radius", C.c_double), ("adapt_engaged", C.c_int), ("unit_e", C.c_int), ("dense_e", C.c_int), ("diag", C.c_int), ("static_hmc", C.c_int), ("int_time", C.c_double)]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "liboracle.so"])


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        _lib = C.CDLL(so)
        _lib.oracle_lp_grad.restype = C.c_double
        _lib.oracle_lp_grad.argtypes = [C.POINTER(ModelT), C.c_void_p, C.c_void_p]
        _lib.oracle_run_chain.restype = C.c_int
        _lib.oracle_run_chain.argtypes = [C.POINTER(ModelT), C.POINTER(ConfigT), C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.POINTER(C.c_long)]
        _lib.oracle_rng_kat.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _lib.oracle_num_outputs.restype = C.c_int
        _lib.oracle_num_outputs.argtypes = [C.POINTER(ModelT)]
        _lib.oracle_write_array.argtypes = [C.POINTER(ModelT), C.c_void_p, C.c_void_p]
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Model:
    """Oracle-side model: name in KIND_BY_NAME + data dict (numpy arrays).  Keeps the arrays alive."""

    def __init__(self, name, data):
        self.name = name
        kind = KIND_BY_NAME[name]
        self._keep = []
        m = ModelT()
        m.kind = kind

        def arr(x, dt):
            a = np.ascontiguousarray(x, dtype=dt)
            self._keep.append(a)
            return a

        if name in ("std_normal", "mock_flat"):
            m.D = int(data.get("D", 3))
        elif name == "bernoulli":
            y = arr(data["y"], np.int32)
            m.D, m.N, m.yi = 1, len(y), _ptr(y)
        elif name == "eight_schools_nc":
            y, s = arr(data["y"], np.float64), arr(data["sigma"], np.float64)
            m.N, m.D, m.yr, m.sigma = len(y), len(y) + 2, _ptr(y), _ptr(s)
        elif name in ("logistic", "poisson_glm"):
            X, y = arr(data["X"], np.float64), arr(data["y"], np.int32)
            m.N, m.K = X.shape
            m.D, m.X, m.yi = m.K + 1, _ptr(X), _ptr(y)
        elif name == "irt_2pl":
            y = arr(data["y"], np.int8)
            m.N, m.K = y.shape
            m.D, m.y8 = 4 + m.N + 2 * m.K, _ptr(y)
        self.c = m
        self.D = m.D

    def lp_grad(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64)
        single = q.ndim == 1
        q2 = q.reshape(-1, self.D)
        lp = np.empty(len(q2))
        g = np.empty_like(q2)
        for i in range(len(q2)):
            lp[i] = lib().oracle_lp_grad(C.byref(self.c), _ptr(q2[i]), _ptr(g[i]))
        return (lp[0], g[0]) if single else (lp, g)

    def num_outputs(self):
        return lib().oracle_num_outputs(C.byref(self.c))

    def write_array(self, q):
        q2 = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, self.D)
        out = np.empty((len(q2), self.num_outputs()))
        for i in range(len(q2)):
            lib().oracle_write_array(C.byref(self.c), _ptr(q2[i]), _ptr(out[i]))
        return out.reshape(q.shape[:-1] + (out.shape[1],))


DEFAULTS = dict(seed=0, chain_id=1, num_warmup=1000, num_samples=1000, num_thin=1, save_warmup=0, stepsize=1.0,
                stepsize_jitter=0.0, max_depth=10, delta=0.8, gamma=0.05, kappa=0.75, t0=10.0, init_buffer=75,
                term_buffer=50, window=25, init_radius=2.0, adapt_engaged=1, unit_e=0, dense_e=0, diag=0, static_hmc=0, int_time=2 * np.pi)


def run_chain(model, q0=None, inv_metric=None, **kw):
    """Run one chain through the C oracle.  Returns dict(draws [rows, 7+D], stepsize, inv_metric, q_init, grad_evals, rc)."""
    cfgd = dict(DEFAULTS)
    cfgd.update(kw)
    cfg = ConfigT(**cfgd)
    D = model.D
    nrow = ((cfg.num_warmup + cfg.num_thin - 1) // cfg.num_thin if cfg.save_warmup else 0) + (
        cfg.num_samples + cfg.num_thin - 1) // cfg.num_thin
    draws = np.zeros((nrow, 7 + D * (3 if cfg.diag else 1)))
    adapt = np.zeros(1 + (D * D if cfg.dense_e else D))
    qinit = np.zeros(D)
    ge = C.c_long(0)
    q0a = None if q0 is None else np.ascontiguousarray(q0, dtype=np.float64)
    ima = None if inv_metric is None else np.ascontiguousarray(inv_metric, dtype=np.float64)
    rc = lib().oracle_run_chain(C.byref(model.c), C.byref(cfg), _ptr(q0a), _ptr(ima), _ptr(draws), _ptr(adapt), _ptr(qinit),
                                C.byref(ge))
    im = adapt[1:].copy().reshape(D, D) if cfg.dense_e else adapt[1:].copy()
    return dict(draws=draws, stepsize=adapt[0], inv_metric=im, q_init=qinit, grad_evals=ge.value, rc=rc)


def rng_kat(seed, chain, kind, n):
    """kind: 0 raw u64, 1 uniform_01, 2 normal, 3 uniform_real(-2,2)."""
    f = np.zeros(n)
    u = np.zeros(n, dtype=np.uint64)
    lib().oracle_rng_kat(seed, chain, kind, n, _ptr(f), _ptr(u))
    return u if kind == 0 else f


# ----------------------------------------------------------------------------------------------------------------
# the compiled, unmodified reference (oracle/_ref/<model> = cmdstan main.cpp + hand-written model class)
# ----------------------------------------------------------------------------------------------------------------
def have_ref(name="bernoulli"):
    return os.path.exists(os.path.join(REF, name))


def write_data_files(name, data, tmpdir):
    """Returns (json_path, env) for the reference executable of model `name`."""
    env = dict(os.environ)
    js = {}
    if name == "std_normal":
        js = {"D": int(data.get("D", 3))}
    elif name == "bernoulli":
        js = {"N": len(data["y"]), "y": [int(v) for v in data["y"]]}
    elif name == "eight_schools_nc":
        js = {"J": len(data["y"]), "y": [float(v) for v in data["y"]], "sigma": [float(v) for v in data["sigma"]]}
    elif name in ("logistic", "poisson_glm"):
        X = np.ascontiguousarray(data["X"], dtype=np.float64)
        y = np.ascontiguousarray(data["y"], dtype=np.int32)
        js = {"N": int(X.shape[0]), "D": int(X.shape[1])}
        binp = os.path.join(tmpdir, name + "_data.bin")
        with open(binp, "wb") as fh:
            fh.write(X.tobytes())
            fh.write(y.tobytes())
        env["B200_ORACLE_DATA_BIN"] = binp
    elif name == "irt_2pl":
        y = np.ascontiguousarray(data["y"], dtype=np.int8)
        js = {"I": int(y.shape[0]), "J": int(y.shape[1])}
        binp = os.path.join(tmpdir, name + "_data.bin")
        with open(binp, "wb") as fh:
            fh.write(y.tobytes())
        env["B200_ORACLE_DATA_BIN"] = binp
    p = os.path.join(tmpdir, name + "_data.json")
    with open(p, "w") as fh:
        json.dump(js, fh)
    return p, env


def read_stan_csv(path):
    """Minimal Stan CSV reader: returns (column names, draws ndarray, stepsize, inv_metric, timing dict)."""
    names, rows, stepsize, metric, timing = None, [], None, None, {}
    with open(path) as fh:
        lines = fh.read().splitlines()
    i = 0
    while i < len(lines):
        ln = lines[i]
        if ln.startswith("#"):
            if ln.startswith("# Step size = "):
                stepsize = float(ln.split("=")[1])
            elif ln.startswith("# Diagonal elements of inverse mass matrix:"):
                metric = np.array([float(v) for v in lines[i + 1][1:].split(",")])
                i += 1
            elif ln.startswith("# Elements of inverse mass matrix:"):  # dense_e_point::write_metric: D rows of D values
                first = [float(v) for v in lines[i + 1][1:].split(",")]
                metric = np.array([first] + [[float(v) for v in lines[i + 1 + k][1:].split(",")] for k in range(1, len(first))])
                i += len(first)
            elif "seconds (" in ln:
                val = float(ln[1:].replace("Elapsed Time:", "").split("seconds")[0])
                timing[ln.split("(")[1].rstrip(")")] = val
        elif ln.strip() == "":
            pass
        elif names is None:
            names = ln.split(",")
        else:
            rows.append([float(v) for v in ln.split(",")])
        i += 1
    return names, np.array(rows), stepsize, metric, timing


def ref_sample(name, data, num_chains=1, seed=0, chain_id=1, num_warmup=1000, num_samples=1000, sig_figs=18, num_threads=None,
               extra_args=(), tmpdir=None, keep=False, save_warmup=False, delta=0.8, max_depth=10, init=None,
               adapt_engaged=True, stepsize=None, metric_file=None, metric=None):
    """Run the unmodified reference `sample` method; returns list (per chain) of read_stan_csv tuples + wall seconds."""
    import time
    own = tmpdir is None
    td = tmpdir or tempfile.mkdtemp(prefix="b200ref_")
    dj, env = write_data_files(name, data, td)
    out = os.path.join(td, "out.csv")
    nt = num_threads or min(num_chains, os.cpu_count() or 1)
    env["STAN_NUM_THREADS"] = str(nt)
    cmd = [os.path.join(REF, name), "sample", f"num_samples={num_samples}", f"num_warmup={num_warmup}",
           f"save_warmup={int(save_warmup)}", "adapt", f"engaged={int(adapt_engaged)}", f"delta={delta}", "algorithm=hmc", "engine=nuts",
           f"max_depth={max_depth}"]
    if metric is not None:
        cmd += [f"metric={metric}"]
    if metric_file is not None:
        cmd += [f"metric_file={metric_file}"]
    if stepsize is not None:
        cmd += [f"stepsize={stepsize}"]
    cmd += [f"num_chains={num_chains}", f"id={chain_id}", "data", f"file={dj}"]
    if init is not None:
        cmd += [f"init={init}"]
    cmd += ["random", f"seed={seed}", "output", f"file={out}", f"sig_figs={sig_figs}", f"num_threads={nt}"]
    cmd += list(extra_args)
    t0 = time.time()
    r = subprocess.run(cmd, env=env, capture_output=True, text=True)
    wall = time.time() - t0
    if r.returncode != 0:
        raise RuntimeError("reference run failed: " + " ".join(cmd) + "\n" + r.stdout[-2000:] + r.stderr[-2000:])
    files = [out] if num_chains == 1 else [os.path.join(td, f"out_{chain_id + i}.csv") for i in range(num_chains)]
    res = [read_stan_csv(f) for f in files]
    return res, wall, files


def ref_log_prob(name, data, q, jacobian=True):
    """Reference `method=log_prob` (C/src/cmdstan/command_helper.hpp:569-600): returns (lp [n], grad [n, D])."""
    td = tempfile.mkdtemp(prefix="b200ref_")
    dj, env = write_data_files(name, data, td)
    q = np.atleast_2d(np.asarray(q, dtype=np.float64))
    up = os.path.join(td, "u.json")
    with open(up, "w") as fh:
        # the reference's JSON reader flattens 2-d arrays column-major while get_uparams_r (command_helper.hpp:548-556)
        # re-reads them row-major; pre-permute so that row i of the result is q[i]
        a = q.ravel().reshape(q.shape, order="F")
        json.dump({"params_r": [[float(v) for v in row] for row in a]}, fh)
    out = os.path.join(td, "lp.csv")
    cmd = [os.path.join(REF, name), "log_prob", f"unconstrained_params={up}", f"jacobian={int(jacobian)}", "data", f"file={dj}",
           "output", f"file={out}", "sig_figs=18"]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("reference log_prob failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
    rows = [ln for ln in open(out).read().splitlines() if ln and not ln.startswith("#")]
    vals = np.array([[float(v) for v in ln.split(",")] for ln in rows[1:]])
    return vals[:, 0], vals[:, 1:]
