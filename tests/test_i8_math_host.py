"""CPU suite: the tensor-core-free arithmetic of the INT8 digit-plane GLM kernel (cmdstan_b200/csrc/glm_i8_math.cuh), built for the
host (tests/host_emul/i8_math_host.cpp): digit extraction folded into the conversion constant + byte transposes == the numpy
restatement's balanced digits (tools/ozaki_i8_model.py) bit for bit; recombination of the INT32 diagonals == exact integers; the
staged four-chain link == fastmath.cuh's logit_term bit for bit (derivative, running product, running sum), cut-offs included.
The kernel around them needs tcgen05 / tensor memory and is validated on hardware (tests/test_gpu_parity.py)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ozaki_i8_model as M  # noqa: E402


@pytest.fixture(scope="module")
def lib():
    so = os.path.join(HERE, "host_emul", "libi8math.so")
    src = os.path.join(HERE, "host_emul", "i8_math_host.cpp")
    deps = [src] + [os.path.join(ROOT, "cmdstan_b200", "csrc", f) for f in ("glm_i8_math.cuh", "fastmath.cuh")]
    if not os.path.exists(so) or max(os.path.getmtime(d) for d in deps) > os.path.getmtime(so):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-o", so, src])
    L = C.CDLL(so)
    L.i8_link_vs_logit_term.restype = C.c_int
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("ns", [5, 6])
def test_digit_planes_equal_the_model(lib, ns):
    F = 8 * ns - 2
    rng = np.random.default_rng(3)
    v = np.concatenate([rng.uniform(-1, 1, 40000), rng.uniform(-1, 1, 4000) * 2.0 ** -rng.integers(1, 45, 4000),
                        [1.0, -1.0, 0.0, -0.0, 0.5, -0.5, 2.0 ** -F, -2.0 ** -F, 1 - 2.0 ** -40, -(1 - 2.0 ** -40), 127.5 / 256, -127.5 / 256]])
    v = np.ascontiguousarray(v[: len(v) // 4 * 4])
    out = np.empty((ns, len(v)), dtype=np.int8)
    lib.i8_planes(C.c_int(ns), _p(v), C.c_int(len(v)), _p(out))
    ref = M.digits(np.rint(v * 2.0 ** F).astype(np.int64), ns)
    for s in range(ns):
        assert np.array_equal(out[s].astype(np.int64), ref[s]), f"plane {s}"


@pytest.mark.parametrize("ns", [5, 6])
def test_recombination_is_exact(lib, ns):
    rng = np.random.default_rng(4)
    # bounds of the Eta diagonals for K = 128 design columns (top digits in [-64, 64], the others in [-128, 127])
    pairs = [sum(1 for i in range(ns) for j in range(ns) if i + j == k) for k in range(ns)]
    bound = [128 * (2 * 64 * 128 + max(p - 2, 0) * 128 * 128) if k else 128 * 64 * 64 for k, p in enumerate(pairs)]
    n = 20000
    diag = np.stack([rng.integers(-b, b + 1, n) for b in bound], axis=1).astype(np.int32)
    diag[:4] = [[b * sgn for b in bound] for sgn in (1, -1, 1, -1)]
    diag[2, ::2] *= -1
    out = np.empty(n, dtype=np.int64)
    lib.i8_combine(C.c_int(ns), _p(np.ascontiguousarray(diag)), C.c_int(n), _p(out))
    ref = [sum(int(diag[i, k]) * 256 ** (ns - 1 - k) for k in range(ns)) for i in range(n)]
    assert [int(x) for x in out] == ref


def test_staged_link_is_logit_term_bit_for_bit(lib):
    rng = np.random.default_rng(5)
    eta = np.concatenate([rng.normal(0, 3, 20000), rng.uniform(-40, 40, 8000), rng.uniform(-800, 800, 2000),
                          [0.0, -0.0, 20.0, -20.0, np.nextafter(20.0, 21), np.nextafter(-20.0, -21), 700.0, -700.0, 1e300, -1e300, np.inf, -np.inf]])
    eta = np.ascontiguousarray(eta[: len(eta) // 4 * 4])
    y = rng.integers(0, 2, len(eta)).astype(np.int32)
    d = np.empty_like(eta)
    assert lib.i8_link_vs_logit_term(_p(eta), _p(y), C.c_int(len(eta)), _p(d)) == 0
    # and the derivative is the logistic one: (2y - 1) / (1 + exp((2y - 1) eta)) inside the cut-offs
    t = (2.0 * y - 1.0) * eta
    m = np.abs(t) < 20
    with np.errstate(over="ignore"):
        ref = ((2.0 * y - 1.0) / (1.0 + np.exp(t.astype(np.longdouble)))).astype(np.float64)
    assert np.max(np.abs(d[m] - ref[m])) < 1e-15
