"""CPU suite: the FP64 tensor-core GLM gradient kernels (cmdstan_b200/csrc/glm_dmma.cu) executed on host threads by the SIMT
stand-in under tests/host_emul/simt (DMMA fragment layouts of the PTX ISA, TMA bulk copies + mbarriers, warp shuffles, the
last-arriver refill protocol), through the product's own launcher.

 * v2 (block-synchronous, the default) and v3 (warp-asynchronous) have run on hardware (profiles/README.md): here they are
   the controls that show the rig reproduces the device semantics;
 * v4 (three X stages, register-shuffle derivative exchange) has NOT run on hardware yet: this is its functional validation;
 * the same sources built with -fsanitize=thread look for shared-memory races in the pipelines, and two deliberately broken
   builds (consumers that do not wait for the bulk copies; a refill issued before the last reader is done) show the race
   detector and the protocol checks have teeth.

A one-off randomised run of the same checks (387 shapes: N 1..700, D 2..112, 1..150 chains, request lists, 1..6 row groups)
found no mismatch in any variant and v4 == v3 bit for bit throughout.

Reference values: numpy restatement of bernoulli_logit_glm_lpmf.hpp:121-142 / poisson_log_glm_lpmf.hpp:117-132 (the sums the
oracle's GLM log-density takes); tolerance 1e-12 relative (gate of the GPU parity tests: 1e-10)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
EMUL = os.path.join(HERE, "host_emul")
SRC = os.path.join(EMUL, "glm_simt.cpp")
SRC2 = os.path.join(EMUL, "glm_simt_stage.cpp")
DEPS = [SRC, SRC2, os.path.join(EMUL, "simt", "cuda_runtime.h")] + [os.path.join(ROOT, "cmdstan_b200", "csrc", f) for f in ("glm_dmma.cu", "glm.cu", "fastmath.cuh", "kernels.h", "models.cuh", "warp.cuh")]
CXX = ["g++", "-std=c++17", "-O1", "-g", "-pthread", "-Wno-tsan", "-I", os.path.join(EMUL, "simt")]


def _stale(out):
    return not os.path.exists(out) or max(os.path.getmtime(d) for d in DEPS) > os.path.getmtime(out)


@pytest.fixture(scope="module")
def lib():
    so = os.path.join(EMUL, "libglm_simt.so")
    if _stale(so):
        subprocess.check_call(CXX + ["-fPIC", "-shared", SRC, SRC2, "-o", so])
    L = C.CDLL(so)
    L.glm_simt_run.argtypes = [C.c_int] * 9 + [C.c_void_p] * 6
    L.glm_simt_grad.argtypes = [C.c_int] * 7 + [C.c_void_p] * 5 + [C.c_int] + [C.c_void_p] * 2
    L.glm_simt_smem_bytes.restype = C.c_size_t
    return L


def _problem(logistic, N, D, Cn, seed):
    rng = np.random.default_rng(seed)
    Dk, ldq = (D + 7) // 8 * 8, (D + 1) // 2 * 2
    X = np.zeros((N, Dk))
    X[:, 0] = 1.0
    X[:, 1:D] = rng.normal(size=(N, D - 1))
    Q = np.zeros((Cn, ldq))
    Q[:, :D] = 0.3 * rng.normal(size=(Cn, D))
    y = (rng.random(N) < 0.5).astype(np.int32) if logistic else rng.poisson(2.0, size=N).astype(np.int32)
    return X, y, Q, Dk, ldq


def _run(lib, variant, logistic, X, y, Q, D, lst, n_list, P):
    N, Dk = X.shape
    Cn, ldq = Q.shape
    Gp = np.full((P, Cn, Dk), np.nan)
    LPp = np.full((P, Cn), np.nan)
    rc = lib.glm_simt_run(variant, int(logistic), N, D, Dk, ldq, Cn, n_list, P, X.ctypes.data, y.ctypes.data, Q.ctypes.data,
                          lst.ctypes.data if lst is not None else None, Gp.ctypes.data, LPp.ctypes.data)
    assert rc == 1
    return Gp, LPp


def _reference(logistic, X, y, Q, D, idx):
    eta = X[:, :D] @ Q[idx, :D].T
    if logistic:
        s = 2.0 * y - 1.0
        t = s[:, None] * eta
        lp = -np.logaddexp(0.0, -t)
        d = s[:, None] / (1.0 + np.exp(t))
    else:
        lp = y[:, None] * eta - np.exp(eta)
        d = y[:, None] - np.exp(eta)
    return lp.sum(0), (X.T @ d).T


# (logistic, N, D, chains, requesting, P, list?)  -- Dk 16/32/64/104/112 select the MB template and 3 vs 2 X stages (v4);
# ragged last row tile, underfilled chain tiles (1, 7, 65 requesting chains), n_it < stages, several chain tiles x row tiles
SHAPES = [
    (1, 300, 13, 70, 70, 2, False),
    (1, 64, 9, 8, 1, 1, True),
    (0, 257, 30, 40, 7, 3, True),
    (1, 500, 51, 90, 65, 2, True),
    (0, 330, 60, 64, 64, 5, False),
    (1, 640, 101, 140, 130, 2, True),
    (0, 450, 110, 70, 70, 3, False),
    (1, 129, 101, 200, 33, 3, True),
]


@pytest.mark.parametrize("variant", [1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "-".join(str(v) for v in s))
def test_partials_against_numpy(lib, variant, shape):
    logistic, N, D, Cn, n_list, P, use_list = shape
    X, y, Q, Dk, _ = _problem(logistic, N, D, Cn, seed=N + D)
    assert lib.glm_simt_smem_bytes(variant, Dk) <= 232448   # the dynamic shared memory a CTA may opt in to on sm_100
    lst = np.random.default_rng(1).permutation(Cn)[:n_list].astype(np.int32) if use_list else None
    Gp, LPp = _run(lib, variant, logistic, X, y, Q, D, lst, n_list, P)
    idx = lst if lst is not None else np.arange(n_list)
    lp_ref, g_ref = _reference(logistic, X, y, Q, D, idx)
    assert np.all(np.isnan(Gp[:, n_list:])) and np.all(np.isnan(LPp[:, n_list:]))   # slots past the list are never written
    assert np.all(np.isfinite(Gp[:, :n_list])) and np.all(np.isfinite(LPp[:, :n_list]))
    LP, G = LPp[:, :n_list].sum(0), Gp[:, :n_list].sum(0)
    assert np.max(np.abs(LP - lp_ref) / np.abs(lp_ref)) < 1e-12
    assert np.max(np.abs(G - g_ref)) / np.max(np.abs(g_ref)) < 1e-12
    assert np.all(G[:, D:] == 0.0)   # padded design columns carry no gradient


@pytest.mark.parametrize("shape", [SHAPES[0], SHAPES[3], SHAPES[5], SHAPES[6]], ids=lambda s: "-".join(str(v) for v in s))
def test_v5_has_the_summation_order_of_v2(lib, shape):
    """glm_dmma.cu, v5 header: two out-of-phase groups of four warps, "results are bit-identical to v2" """
    logistic, N, D, Cn, n_list, P, use_list = shape
    X, y, Q, _, _ = _problem(logistic, N, D, Cn, seed=6)
    lst = np.arange(n_list, dtype=np.int32)[::-1].copy() if use_list else None
    a = _run(lib, 1, logistic, X, y, Q, D, lst, n_list, P)
    b = _run(lib, 4, logistic, X, y, Q, D, lst, n_list, P)
    assert np.array_equal(a[0][:, :n_list], b[0][:, :n_list]) and np.array_equal(a[1][:, :n_list], b[1][:, :n_list])


@pytest.mark.parametrize("shape", [SHAPES[0], SHAPES[3], SHAPES[5], SHAPES[6]], ids=lambda s: "-".join(str(v) for v in s))
def test_v4_has_the_summation_order_of_v3(lib, shape):
    """glm_dmma.cu, v4 header: "Summation order per chain is that of v3" -- the partials agree bit for bit"""
    logistic, N, D, Cn, n_list, P, use_list = shape
    X, y, Q, _, _ = _problem(logistic, N, D, Cn, seed=5)
    lst = np.arange(n_list, dtype=np.int32)[::-1].copy() if use_list else None
    a = _run(lib, 2, logistic, X, y, Q, D, lst, n_list, P)
    b = _run(lib, 3, logistic, X, y, Q, D, lst, n_list, P)
    assert np.array_equal(a[0][:, :n_list], b[0][:, :n_list]) and np.array_equal(a[1][:, :n_list], b[1][:, :n_list])


@pytest.mark.parametrize("variant", [1, 2, 3, 4, 5, 6])
def test_a_chain_does_not_depend_on_its_batch(lib, variant):
    """kernels.h: results for a chain depend only on that chain's q -- whichever other chains request a gradient in the same
    launch and whichever slot of the chain tile it lands in (hmc_nuts_diag_e_adapt_parallel_match_test.cpp:73-153)"""
    logistic, N, D, Cn, P = 1, 400, 21, 100, 3
    X, y, Q, _, _ = _problem(logistic, N, D, Cn, seed=11)
    full = np.arange(Cn, dtype=np.int32)
    Gf, Lf = _run(lib, variant, logistic, X, y, Q, D, full, Cn, P)
    sub = np.array([93, 5, 64, 17, 70, 2, 99, 31, 8, 50, 63], dtype=np.int32)
    Gs, Ls = _run(lib, variant, logistic, X, y, Q, D, sub, len(sub), P)
    assert np.array_equal(Gs[:, :len(sub)], Gf[:, sub]) and np.array_equal(Ls[:, :len(sub)], Lf[:, sub])


# ---- the whole gradient stage (partial kernel of each variant, reduction over row groups, priors and signs) vs the oracle -------
@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("name,N,K", [("logistic", 300, 12), ("poisson_glm", 257, 30), ("logistic", 200, 100)])
def test_gradient_stage_against_the_oracle(lib, variant, name, N, K):
    """launch_glm_grad (kernels.h): zV = -lp, G = -grad lp with lp = log_prob_propto_jacobian of the oracle's GLM
    (bernoulli_logit_glm / poisson_log_glm + normal priors), for requesting chains only; 1e-12 (GPU gate: 1e-10)"""
    from oracle import oracle as O
    import cases
    data = cases.logistic_data(N, K, seed=3) if name == "logistic" else cases.poisson_data(N, K, seed=3)
    D, Cn = K + 1, 70
    ldq = (D + 3) & ~3
    Dk = (D + 7) & ~7
    X = np.zeros((N, Dk))
    X[:, 0] = 1.0
    X[:, 1:D] = data["X"]
    y = np.ascontiguousarray(data["y"], dtype=np.int32)
    Q = np.zeros((Cn, ldq))
    Q[:, :D] = 0.2 * np.random.default_rng(K).normal(size=(Cn, D))
    req = np.array(sorted(np.random.default_rng(2).permutation(Cn)[:41]), dtype=np.int32)
    active = np.zeros(Cn, dtype=np.int32)
    active[req] = 1
    G, zV = np.full((Cn, ldq), np.nan), np.full(Cn, np.nan)
    n = lib.glm_simt_grad(variant, int(name == "logistic"), N, D, ldq, Cn, 5, X.ctypes.data, y.ctypes.data, Q.ctypes.data, active.ctypes.data,
                          req.ctypes.data, len(req), G.ctypes.data, zV.ctypes.data)
    assert n == 3
    lp, g = O.Model(name, data).lp_grad(Q[req][:, :D])
    assert np.max(np.abs(zV[req] + lp) / np.abs(lp)) < 1e-12
    assert np.max(np.abs(G[req][:, :D] + g)) / np.max(np.abs(g)) < 1e-12
    idle = np.setdiff1d(np.arange(Cn), req)
    assert np.all(np.isnan(zV[idle])) and np.all(np.isnan(G[idle]))   # chains that did not ask are not touched


@pytest.mark.parametrize("variant", [0, 1, 3, 4, 5, 6])
def test_gradient_stage_edge_values(lib, variant):
    """the +-20 cut-offs of bernoulli_logit_glm (linear predictors of +-60), and poisson_log_glm with exp(eta) overflowing:
    the reference throws a domain error there (non-finite sum of the derivative) => rejected proposal, V = +inf"""
    from oracle import oracle as O
    import cases
    N, K, Cn = 130, 6, 9
    D = K + 1
    ldq, Dk = (D + 3) & ~3, (D + 7) & ~7
    for name in ("logistic", "poisson_glm"):
        data = cases.logistic_data(N, K, seed=4) if name == "logistic" else cases.poisson_data(N, K, seed=4)
        data = dict(data, X=data["X"] * (10.0 if name == "logistic" else 1.0))
        X = np.zeros((N, Dk))
        X[:, 0] = 1.0
        X[:, 1:D] = data["X"]
        y = np.ascontiguousarray(data["y"], dtype=np.int32)
        Q = np.zeros((Cn, ldq))
        Q[:, :D] = np.random.default_rng(1).normal(size=(Cn, D)) * np.linspace(0.1, 4.0, Cn)[:, None]
        if name == "poisson_glm":
            Q[Cn - 1, 0] = 800.0    # exp overflows for every row
            Q[Cn - 2, 0] = 600.0    # finite exp, huge derivative
        G, zV = np.full((Cn, ldq), np.nan), np.full(Cn, np.nan)
        lib.glm_simt_grad(variant, int(name == "logistic"), N, D, ldq, Cn, 3, X.ctypes.data, y.ctypes.data, Q.ctypes.data, None, None, 0, G.ctypes.data, zV.ctypes.data)
        lp, g = O.Model(name, data).lp_grad(Q[:, :D])
        if name == "logistic":
            assert np.abs(X[:, :D] @ Q[:, :D].T).max() > 60.0
        ok = np.isfinite(lp)
        assert ok.sum() >= Cn - 1
        assert np.array_equal(np.isposinf(zV), ~ok)         # V = +inf exactly where the oracle's lp is -inf
        assert np.max(np.abs(zV[ok] + lp[ok]) / np.abs(lp[ok])) < 1e-12
        assert np.max(np.abs(G[ok][:, :D] + g[ok]) / np.max(np.abs(g[ok]), axis=1, keepdims=True)) < 1e-12


# ---- ThreadSanitizer ---------------------------------------------------------------------------------------------------
def _tsan_exe(name, defines=()):
    exe = os.path.join(EMUL, name + ".exe")
    if _stale(exe):
        cxx = [f for f in CXX if f not in ("-O1", "-g")] + ["-O0"] if defines else CXX   # the broken builds need no line info: 3x faster to compile
        r = subprocess.run(cxx + ["-fsanitize=thread", "-DGLM_SIMT_MAIN"] + ["-D" + d for d in defines] + [SRC, SRC2, "-o", exe], capture_output=True, text=True)
        if r.returncode != 0:
            pytest.skip("no ThreadSanitizer tool chain: " + r.stderr[-300:])
    return exe


def _tsan_run(exe, args):
    env = dict(os.environ, TSAN_OPTIONS="exitcode=66 halt_on_error=0")
    r = subprocess.run([exe] + [str(a) for a in args], capture_output=True, text=True, env=env, timeout=900)
    if "FATAL: ThreadSanitizer" in r.stderr and "data race" not in r.stderr:
        pytest.skip("ThreadSanitizer cannot run here: " + r.stderr[-300:])
    return r


# variant logistic N D chains requesting P
TSAN_CASES = [(0, 1, 300, 51, 70, 70, 3), (1, 1, 700, 51, 200, 77, 3), (2, 1, 640, 101, 140, 130, 2), (3, 1, 640, 101, 140, 130, 2), (3, 0, 900, 110, 70, 70, 3), (3, 1, 200, 13, 70, 70, 3),
              (4, 1, 640, 101, 140, 130, 2), (4, 0, 900, 110, 70, 40, 3), (4, 1, 200, 13, 70, 70, 3),
              (5, 1, 640, 101, 140, 130, 2), (5, 0, 900, 110, 70, 40, 3), (5, 1, 200, 13, 70, 70, 3),
              (6, 1, 640, 101, 140, 130, 2), (6, 0, 900, 110, 70, 40, 3), (6, 1, 200, 13, 70, 70, 3)]


@pytest.mark.parametrize("case", TSAN_CASES, ids=lambda c: "-".join(str(v) for v in c))
def test_no_shared_memory_race(case):
    r = _tsan_run(_tsan_exe("glm_simt_tsan"), case)
    assert r.returncode == 0 and "ThreadSanitizer" not in r.stderr and "\nsimt: " not in "\n" + r.stderr, r.stderr[-3000:]
    assert r.stdout.startswith("lp checksum")


@pytest.mark.parametrize("mutant", ["SIMT_MUTANT_SKIP_MBAR_WAIT", "SIMT_MUTANT_EARLY_REFILL"])
def test_the_race_detector_sees_a_broken_pipeline(mutant):
    r = _tsan_run(_tsan_exe("glm_simt_tsan_" + mutant.lower(), [mutant]), (3, 1, 320, 13, 70, 70, 1))
    assert r.returncode != 0 and ("data race" in r.stderr or "simt:" in r.stderr)
