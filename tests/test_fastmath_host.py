"""CPU suite: accuracy of the branch-free link functions the GLM and IRT gradient kernels evaluate per observation
(cmdstan_b200/csrc/fastmath.cuh), built for the host (tests/host_emul/fastmath_host.cpp) and compared with 80-bit references:
fast_exp <= 4e-16 relative, the logistic terms <= 1e-15 (measured 7e-16); the gate on sums of 1e5..1e7 such terms is 1e-10 (north_star)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def lib():
    so = os.path.join(HERE, "host_emul", "libfastmath.so")
    src = os.path.join(HERE, "host_emul", "fastmath_host.cpp")
    hdr = os.path.join(ROOT, "cmdstan_b200", "csrc", "fastmath.cuh")
    if not os.path.exists(so) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(so):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-o", so, src])
    return C.CDLL(so)


def _link(lib, link, eta, yd):
    eta, yd = np.ascontiguousarray(eta, dtype=np.float64), np.ascontiguousarray(yd, dtype=np.float64)
    lp, d = np.empty_like(eta), np.empty_like(eta)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.fastmath_link(C.c_int(link), p(eta), p(yd), C.c_int(len(eta)), p(lp), p(d))
    return lp, d


def test_fast_exp_accuracy(lib):
    x = np.random.default_rng(0).uniform(-700, 700, size=400000)
    out = np.empty_like(x)
    lib.fastmath_exp(x.ctypes.data_as(C.c_void_p), C.c_int(len(x)), out.ctypes.data_as(C.c_void_p))
    ref = np.exp(x.astype(np.longdouble))
    assert float(np.max(np.abs(out.astype(np.longdouble) - ref) / ref)) < 4e-16


def test_logistic_link_accuracy_and_cutoffs(lib):
    """bernoulli_logit_glm_lpmf.hpp:121-126, :137-142: lp = -log1p(exp(-t)), d = sign * e / (e + 1) with e = exp(-t), t = sign * eta;
    t > 20: lp = d = -exp(-t) (the reference does NOT multiply the derivative by the sign there); t < -20: lp = t, d = sign."""
    rng = np.random.default_rng(1)
    eta = np.concatenate([rng.uniform(-19.9, 19.9, size=300000), rng.normal(0, 1e-3, size=20000), rng.normal(0, 1e-9, size=1000)])
    yd = rng.choice([-1.0, 1.0], size=len(eta))
    lp, d = _link(lib, 0, eta, yd)
    t = (yd * eta).astype(np.longdouble)
    e = np.exp(-t)
    lp_ref, d_ref = -np.log1p(e), yd.astype(np.longdouble) * e / (e + 1)
    assert float(np.max(np.abs(lp - lp_ref) / np.abs(lp_ref))) < 1e-15
    assert float(np.max(np.abs(d - d_ref) / np.abs(d_ref))) < 1e-15
    # cut-offs and special values
    eta = np.array([20.0, 20.0000001, 25.0, 700.0, 800.0, -20.0, -20.0000001, -35.0, -800.0, 0.0, -0.0, np.inf, -np.inf, np.nan])
    for sign in (1.0, -1.0):
        yd = np.full(len(eta), sign)
        lp, d = _link(lib, 0, eta, yd)
        for i, x in enumerate(eta):
            tt = sign * x
            if np.isnan(tt):
                assert np.isnan(lp[i]) and np.isnan(d[i])
            elif tt > 20:
                want = -np.exp(-min(tt, 700.0)) if np.isfinite(tt) else -np.exp(-700.0)
                assert lp[i] == d[i] and abs(lp[i] - want) <= 1e-15 * abs(want) + 1e-320
            elif tt < -20:
                assert lp[i] == tt and d[i] == sign
            else:
                ee = np.exp(np.longdouble(-tt))
                assert abs(lp[i] - float(-np.log1p(ee))) <= 1e-15 * abs(float(np.log1p(ee)))
                assert abs(d[i] - float(sign * ee / (ee + 1))) <= 1e-15


def test_poisson_link_accuracy_and_overflow(lib):
    """poisson_log_glm_lpmf.hpp:117-132: lp = y eta - exp(eta), d = y - exp(eta); exp overflows to +inf above 709.78"""
    rng = np.random.default_rng(2)
    eta = rng.uniform(-30, 30, size=200000)
    y = rng.integers(0, 50, size=len(eta)).astype(np.float64)
    lp, d = _link(lib, 1, eta, y)
    ee = np.exp(eta.astype(np.longdouble))
    d_ref = y.astype(np.longdouble) - ee
    assert float(np.max(np.abs(d - d_ref) / np.maximum(y, ee))) < 6e-16
    lp_ref = y.astype(np.longdouble) * eta.astype(np.longdouble) - ee
    assert float(np.max(np.abs(lp - lp_ref) / np.maximum(np.abs(lp_ref), ee))) < 6e-16
    lp, d = _link(lib, 1, np.array([709.0, 710.0, 1e6, -1e6, np.nan]), np.array([3.0, 3.0, 3.0, 3.0, 3.0]))
    assert np.isfinite(lp[0]) and lp[1] == -np.inf and d[1] == -np.inf and lp[2] == -np.inf
    assert lp[3] == -3e6 and d[3] == 3.0 and np.isnan(lp[4]) and np.isnan(d[4])
