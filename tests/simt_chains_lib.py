"""TEST-ONLY ctypes wrapper for tests/host_emul/libchains_simt.so: the engine's chain kernels (fused, shared-memory staged,
lock-step) run on host threads with real 32-lane warps (tests/host_emul/simt)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
EMUL = os.path.join(HERE, "host_emul")
SRC = os.path.join(EMUL, "chains_simt.cpp")
DEPS = [SRC, os.path.join(EMUL, "simt", "cuda_runtime.h")] + [os.path.join(ROOT, "cmdstan_b200", "csrc", f)
                                                               for f in ("chains.cuh", "nuts.cuh", "rng.cuh", "models.cuh", "warp.cuh", "host_seed.h")]
CXX = ["g++", "-std=c++17", "-O1", "-g", "-pthread", "-ffp-contract=off", "-Wno-tsan", "-Wno-unknown-pragmas", "-I", os.path.join(EMUL, "simt")]
_lib = None


def stale(out):
    return not os.path.exists(out) or max(os.path.getmtime(d) for d in DEPS) > os.path.getmtime(out)


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(EMUL, "libchains_simt.so")
        if stale(so):
            subprocess.check_call(CXX + ["-fPIC", "-shared", SRC, "-o", so])
        _lib = C.CDLL(so)
        _lib.simt_run_chains.restype = C.c_longlong
    return _lib


def run_chains(mode, n_chains, kind, D, data, seed=0, chain_id=1, num_warmup=1000, num_samples=1000, num_thin=1, save_warmup=0, stepsize=1.0,
               stepsize_jitter=0.0, max_depth=10, delta=0.8, gamma=0.05, kappa=0.75, t0=10.0, init_buffer=75, term_buffer=50, window=25,
               init_radius=2.0, adapt_engaged=1, q0=None, inv_metric=None, unit_e=0, dense_e=0, diag=0, static_hmc=0, int_time=2 * np.pi):
    """q0 / inv_metric: one chain's worth (replicated) or [n_chains][...]"""
    yr = np.ascontiguousarray(data.get("y", [0.0]), dtype=np.float64)
    sg = np.ascontiguousarray(data.get("sigma", [1.0]), dtype=np.float64)
    N = len(yr) if kind in (1, 2) else 0
    sum_y = int(np.sum(data["y"])) if kind == 1 else 0
    nrow = ((num_warmup + num_thin - 1) // num_thin if save_warmup else 0) + (num_samples + num_thin - 1) // num_thin
    draws = np.zeros((n_chains, nrow, 7 + D * (3 if diag else 1)))
    ao = 1 + D + (D * D if dense_e else 0)
    adapt = np.zeros((n_chains, ao))
    ng = np.zeros(n_chains, dtype=np.int64)
    err = np.zeros(n_chains, dtype=np.int32)

    def per_chain(a, shape):
        if a is None:
            return None
        a = np.asarray(a, dtype=np.float64)
        if a.shape == shape:
            a = np.broadcast_to(a, (n_chains,) + shape)
        return np.ascontiguousarray(a)
    q0a = per_chain(q0, (D,))
    ima = per_chain(inv_metric, (D, D) if dense_e else (D,))
    vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    launches = lib().simt_run_chains(
        C.c_int(mode), C.c_int(n_chains), C.c_int(kind), C.c_int(D), C.c_int(N), vp(yr), vp(sg), C.c_int(sum_y), C.c_uint(seed), C.c_uint(chain_id),
        C.c_int(num_warmup), C.c_int(num_samples), C.c_int(num_thin), C.c_int(save_warmup), C.c_double(stepsize), C.c_double(stepsize_jitter),
        C.c_int(max_depth), C.c_double(delta), C.c_double(gamma), C.c_double(kappa), C.c_double(t0), C.c_uint(init_buffer), C.c_uint(term_buffer),
        C.c_uint(window), C.c_double(init_radius), C.c_int(adapt_engaged), vp(q0a), vp(ima), vp(draws), vp(adapt), vp(ng), vp(err), C.c_int(unit_e),
        C.c_int(dense_e), C.c_int(diag), C.c_int(static_hmc), C.c_double(int_time))
    im = adapt[:, 1:1 + D * D].reshape(n_chains, D, D).copy() if dense_e else adapt[:, 1:1 + D].copy()
    return dict(draws=draws, stepsize=adapt[:, 0].copy(), inv_metric=im, n_grad=ng, rc=err, launches=launches)
