"""TEST-ONLY ctypes wrapper for tests/host_emul/libemul.so (device state machine compiled for the host, one-lane warp)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
_lib = None


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(HERE, "host_emul", "libemul.so")
        src = os.path.join(HERE, "host_emul", "emul.cpp")
        deps = [src] + [os.path.join(ROOT, "cmdstan_b200", "csrc", f) for f in ("nuts.cuh", "rng.cuh", "models.cuh", "warp.cuh", "host_seed.h")]
        if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-DB200_HOST_EMUL", "-shared", "-o", so, src])
        _lib = C.CDLL(so)
        _lib.emul_run_chain.restype = C.c_int
    return _lib


def run_chain(kind, D, data, seed=0, chain_id=1, num_warmup=1000, num_samples=1000, num_thin=1, save_warmup=0, stepsize=1.0,
              stepsize_jitter=0.0, max_depth=10, delta=0.8, gamma=0.05, kappa=0.75, t0=10.0, init_buffer=75, term_buffer=50,
              window=25, init_radius=2.0, adapt_engaged=1, q0=None, inv_metric=None, unit_e=0, dense_e=0, diag=0, static_hmc=0, int_time=2 * np.pi):
    yr = np.ascontiguousarray(data.get("y", [0.0]), dtype=np.float64)
    sg = np.ascontiguousarray(data.get("sigma", [1.0]), dtype=np.float64)
    N = len(yr) if kind in (1, 2) else 0
    sum_y = int(np.sum(data["y"])) if kind == 1 else 0
    nrow = ((num_warmup + num_thin - 1) // num_thin if save_warmup else 0) + (num_samples + num_thin - 1) // num_thin
    draws = np.zeros((nrow, 7 + D * (3 if diag else 1)))
    adapt = np.zeros(1 + (D * D if dense_e else D))
    ng = C.c_longlong(0)
    p = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(C.c_void_p)
    q0a = None if q0 is None else np.ascontiguousarray(q0, dtype=np.float64)
    ima = None if inv_metric is None else np.ascontiguousarray(inv_metric, dtype=np.float64)
    rc = lib().emul_run_chain(C.c_int(kind), C.c_int(D), C.c_int(N), yr.ctypes.data_as(C.c_void_p), sg.ctypes.data_as(C.c_void_p),
                              C.c_int(sum_y), C.c_uint(seed), C.c_uint(chain_id), C.c_int(num_warmup), C.c_int(num_samples),
                              C.c_int(num_thin), C.c_int(save_warmup), C.c_double(stepsize), C.c_double(stepsize_jitter),
                              C.c_int(max_depth), C.c_double(delta), C.c_double(gamma), C.c_double(kappa), C.c_double(t0),
                              C.c_uint(init_buffer), C.c_uint(term_buffer), C.c_uint(window), C.c_double(init_radius),
                              C.c_int(adapt_engaged), None if q0a is None else q0a.ctypes.data_as(C.c_void_p),
                              None if ima is None else ima.ctypes.data_as(C.c_void_p), draws.ctypes.data_as(C.c_void_p),
                              adapt.ctypes.data_as(C.c_void_p), C.byref(ng), C.c_int(unit_e), C.c_int(dense_e), C.c_int(diag), C.c_int(static_hmc), C.c_double(int_time))
    im = adapt[1:].copy().reshape(D, D) if dense_e else adapt[1:].copy()
    return dict(draws=draws, stepsize=adapt[0], inv_metric=im, n_grad=ng.value, rc=rc)
