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


# ---- round 2: table-driven exp, cubic reciprocal, product form of the logistic log-likelihood ---------------------------
def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_fast_exp_table_accuracy(lib):
    x = np.concatenate([np.random.default_rng(3).uniform(-700, 700, size=400000), np.array([0.0, -0.0, -700.0, 700.0, 1e-300, -1e-17])])
    out = np.empty_like(x)
    lib.fastmath_exp_t(_p(x), C.c_int(len(x)), _p(out))
    ref = np.exp(x.astype(np.longdouble))
    assert float(np.max(np.abs(out.astype(np.longdouble) - ref) / ref)) < 3.5e-16


def test_fast_rcp3_accuracy(lib):
    x = np.random.default_rng(4).uniform(1.0, 2.0, size=200000)
    out = np.empty_like(x)
    lib.fastmath_rcp3(_p(x), C.c_int(len(x)), _p(out))
    assert float(np.max(np.abs(out.astype(np.longdouble) * x.astype(np.longdouble) - 1))) < 2.5e-16


def test_logit_product_form_terms_and_cutoffs(lib):
    """per observation: derivative to 2.5e-16 absolute + 1e-15 relative; the term itself to 2.5e-16 ABSOLUTE + 1e-15 relative (a
    product in (1, 2] carries an absolute, not a relative, rounding error in its logarithm: what matters is the sum, next test)."""
    rng = np.random.default_rng(5)
    eta = np.concatenate([rng.uniform(-19.9, 19.9, size=300000), rng.normal(0, 1e-3, size=20000), rng.normal(0, 1e-9, size=1000)])
    yd = rng.choice([-1.0, 1.0], size=len(eta))
    lp, d = np.empty_like(eta), np.empty_like(eta)
    lib.fastmath_logit_terms(_p(eta), _p(yd), C.c_int(len(eta)), _p(lp), _p(d))
    t = (yd * eta).astype(np.longdouble)
    e = np.exp(-t)
    lp_ref, d_ref = -np.log1p(e), yd.astype(np.longdouble) * e / (e + 1)
    assert float(np.max(np.abs(d - d_ref) - 1e-15 * np.abs(d_ref))) < 2.5e-16
    assert float(np.max(np.abs(lp - lp_ref) - 1e-15 * np.abs(lp_ref))) < 2.5e-16
    # cut-offs (bernoulli_logit_glm_lpmf.hpp:121-126, :137-142) and special values
    eta = np.array([20.0, 20.00001, 25.0, 700.0, 800.0, 1e300, -20.0, -20.00001, -35.0, -800.0, -1e300, 0.0, -0.0, np.inf, -np.inf])
    for sign in (1.0, -1.0):
        yd = np.full(len(eta), sign)
        lp, d = np.empty_like(eta), np.empty_like(eta)
        lib.fastmath_logit_terms(_p(eta), _p(yd), C.c_int(len(eta)), _p(lp), _p(d))
        for i, x in enumerate(eta):
            tt = sign * x
            if tt > 20:      # term -exp(-t), derivative -exp(-t) whatever the sign
                want = -float(np.exp(np.longdouble(-tt)))
                assert abs(lp[i] - want) <= 1e-15 * abs(want) + 2.5e-16 and abs(d[i] - want) <= 1e-15 * abs(want) + 2.5e-16, (tt, lp[i], d[i], want)
            elif tt < -20:   # exactly t and the sign
                assert lp[i] == tt and d[i] == sign, (tt, lp[i], d[i])
            else:
                ee = np.exp(np.longdouble(-tt))
                assert abs(lp[i] - float(-np.log1p(ee))) <= 1e-15 * abs(float(np.log1p(ee))) + 2.5e-16
                assert abs(d[i] - float(sign * ee / (ee + 1))) <= 1e-15


@pytest.mark.parametrize("scale,fold_every,n_pad", [(1.0, 7, 0), (3.0, 64, 1000), (8.0, 900, 37), (0.01, 333, 5)])
def test_logit_product_form_sums(lib, scale, fold_every, n_pad):
    """what the kernels compute: sums of 1e5 terms through one accumulator (exponent folded every so often), padding observations
    mixed in and removed exactly: relative error of the sum <= 1e-14 (measured ~1e-16; gate on lp: 1e-10)."""
    rng = np.random.default_rng(6)
    n = 100000
    eta = rng.normal(0, scale, size=n)
    eta[::997] *= 40.0           # a few terms beyond both cut-offs
    yd = rng.choice([-1.0, 1.0], size=n)
    d = np.empty_like(eta)
    lib.fastmath_logit_sum.restype = C.c_double
    got = lib.fastmath_logit_sum(_p(eta), _p(yd), C.c_int(n), C.c_int(fold_every), C.c_int(n_pad), _p(d))
    t = (yd * eta).astype(np.longdouble)
    e = np.exp(-np.minimum(np.abs(t), 700))
    term = np.where(t > 20, -np.exp(-np.minimum(t, 700)), np.where(t < -20, t, -(np.where(t < 0, -t, 0) + np.log1p(e))))
    ref = float(np.sum(term))
    assert abs(got - ref) <= 1e-14 * abs(ref)
