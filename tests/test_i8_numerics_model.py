"""CPU suite: the arithmetic of the INT8 digit-plane GLM kernel (cmdstan_b200/csrc/glm_i8.cu) restated exactly in numpy
(tools/ozaki_i8_model.py: balanced base-256 digits, per-diagonal integer accumulation, dropped diagonals, power-of-two scales)
against an extended-precision evaluation -- the error bounds DESIGN.md 3.6 states and the GPU tests assert on hardware."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tools"))
sys.path.insert(0, HERE)
import cases  # noqa: E402
import ozaki_i8_model as M  # noqa: E402


@pytest.mark.parametrize("ns", [5, 6])
def test_digits_reconstruct_and_fit_int8(ns):
    rng = np.random.default_rng(1)
    F = 8 * ns - 2
    v = np.concatenate([rng.uniform(-1, 1, 5000), [1.0, -1.0, 0.0, 2.0 ** -F, -2.0 ** -F, 0.5, 127.5 / 256]])
    V = np.rint(v * 2.0 ** F).astype(np.int64)
    d = M.digits(V, ns)
    assert all(np.all(x >= -128) and np.all(x <= 127) for x in d) and np.all(np.abs(d[0]) <= 64)
    assert np.array_equal(sum(x * (1 << (8 * (ns - 1 - i))) for i, x in enumerate(d)), V)


@pytest.mark.parametrize("N,K,nq", [(1500, 20, 12), (400, 101, 8)])
def test_model_error_bounds(N, K, nq):
    data = cases.logistic_data(N, K)
    X = np.concatenate([np.ones((N, 1)), data["X"]], axis=1)
    Q = np.random.default_rng(5).uniform(-1.5, 1.5, size=(nq, K + 1))
    lp0, g0 = M.lp_grad_exact(X, data["y"], Q)
    gmax = np.max(np.abs(g0))
    lp6, g6 = M.lp_grad_model(X, data["y"], Q, 6)
    lp5, g5 = M.lp_grad_model(X, data["y"], Q, 5)
    # six planes (the default): what the kernel delivers on hardware (2.5e-14 / 8e-14) with a margin for the shapes here
    assert M.rel(lp6, lp0, 1e-12) < 2e-13 and np.max(np.abs(g6 - g0)) / gmax < 5e-13
    # five planes (B200NUTS_I8_SLICES=5): still inside north_star's 1e-10
    assert M.rel(lp5, lp0, 1e-12) < 5e-11 and np.max(np.abs(g5 - g0)) / gmax < 1e-10
    # the gate the GPU tests apply to the kernel (tests/test_gpu_parity.py::_rel_grad)
    floor = 1e-2 * np.max(np.abs(g0), axis=-1, keepdims=True)
    assert np.max(np.abs(g6 - g0) / np.maximum(floor, np.abs(g0))) < 1e-10
