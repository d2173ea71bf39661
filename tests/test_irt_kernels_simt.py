"""CPU suite: the IRT 2PL kernels (cmdstan_b200/csrc/irt.cu) executed on host threads by the SIMT stand-in under
tests/host_emul/simt, through the product's launchers.

 * the throughput gradient kernel in its two builds (one CTA per SM, two CTAs per SM) against the oracle's log-density and
   gradient (tolerance 1e-12; gate of the GPU parity tests: 1e-10), with request lists / active flags, for the items-per-lane
   templates 2, 4, 13; the two builds perform the same operations in the same order (equal bit for bit);
 * the latency kernel (irt_solo_kernel, here in its one-CTA-per-chain form; the eight-CTA cluster form shares everything but the
   distributed-shared-memory gather and is covered on hardware): the person-block arithmetic, the CTA-cooperative leaf step
   (solo_leaf.cuh) and the generic advance() fallback: chains handed to it after any number of lock-steps must finish with the
   draws of the lock-step engine bit for bit; ThreadSanitizer sees no race between its phases."""
# A one-off randomised run of the gradient checks (917 shapes: 1..129 persons, 1..128 items -- every items-per-lane template --,
# 1..4 chains, parameter scales 0.3 / 1 / 3) found no mismatch with the oracle and no difference between the three builds.
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import cases
from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
EMUL = os.path.join(HERE, "host_emul")
SRC = os.path.join(EMUL, "irt_simt.cpp")
DEPS = [SRC, os.path.join(EMUL, "simt", "cuda_runtime.h")] + [os.path.join(ROOT, "cmdstan_b200", "csrc", f) for f in
                                                               ("irt.cu", "advance_cta.cu", "solo_leaf.cuh", "irt_body.cuh", "chains.cuh", "nuts.cuh", "rng.cuh", "fastmath.cuh", "warp.cuh", "models.cuh")]
CXX = ["g++", "-std=c++17", "-O1", "-g", "-pthread", "-ffp-contract=off", "-Wno-tsan", "-Wno-unknown-pragmas", "-I", os.path.join(EMUL, "simt")]
vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)


def _stale(out):
    return not os.path.exists(out) or max(os.path.getmtime(d) for d in DEPS) > os.path.getmtime(out)


@pytest.fixture(scope="module")
def lib():
    so = os.path.join(EMUL, "libirt_simt.so")
    if _stale(so):
        subprocess.check_call(CXX + ["-fPIC", "-shared", SRC, "-o", so])
    L = C.CDLL(so)
    L.irt_simt_run.restype = C.c_longlong
    return L


def _grad(lib, form, I, J, Cn, active=None, lst=None):
    data = cases.irt_data(I, J, seed=5)
    y = np.ascontiguousarray(data["y"], dtype=np.int8)
    D = 4 + I + 2 * J
    ld = (D + 3) & ~3
    Q = np.zeros((Cn, ld))
    Q[:, :D] = 0.5 * np.random.default_rng(I + J).normal(size=(Cn, D))
    G, zV = np.full((Cn, ld), np.nan), np.full(Cn, np.nan)
    a = None if active is None else np.ascontiguousarray(active, dtype=np.int32)
    l = None if lst is None else np.ascontiguousarray(lst, dtype=np.int32)
    assert lib.irt_simt_grad(form, I, J, Cn, ld, vp(y), vp(Q), vp(a), vp(l), 0 if l is None else len(l), vp(G), vp(zV)) == 1
    lp, g = O.Model("irt_2pl", data).lp_grad(Q[:, :D])
    return G[:, :D], zV, -g, -lp


@pytest.mark.parametrize("shape", [(37, 11), (20, 100), (64, 29), (5, 1)], ids=lambda s: "%dx%d" % s)
def test_gradient_kernel_builds_against_the_oracle(lib, shape):
    I, J = shape
    out = []
    for form in (1, 2):
        G, zV, g_ref, v_ref = _grad(lib, form, I, J, 3)
        assert np.max(np.abs(G - g_ref)) / np.max(np.abs(g_ref)) < 1e-12 and np.max(np.abs(zV - v_ref) / np.abs(v_ref)) < 1e-12
        out.append((G, zV))
    for G, zV in out[1:]:
        assert np.array_equal(G, out[0][0]) and np.array_equal(zV, out[0][1])


@pytest.mark.parametrize("form", [1, 2])
def test_request_list_and_active_flags(lib, form):
    I, J, Cn = 23, 9, 6
    full = _grad(lib, form, I, J, Cn)
    G, zV, _, _ = _grad(lib, form, I, J, Cn, active=[0, 1, 0, 0, 1, 1])
    for c, on in enumerate([0, 1, 0, 0, 1, 1]):
        if on:
            assert np.array_equal(G[c], full[0][c]) and zV[c] == full[1][c]
        else:
            assert np.all(np.isnan(G[c])) and np.isnan(zV[c])     # chains that did not ask are not touched
    G, zV, _, _ = _grad(lib, form, I, J, Cn, lst=[4, 0, 3])
    for c in range(Cn):
        if c in (4, 0, 3):
            assert np.array_equal(G[c], full[0][c]) and zV[c] == full[1][c]
        else:
            assert np.all(np.isnan(G[c])) and np.isnan(zV[c])


def _run(lib, Cn, I, J, nw, ns, max_depth, tail_after, seed=3):
    y = np.ascontiguousarray(cases.irt_data(I, J, seed=5)["y"], dtype=np.int8)
    D = 4 + I + 2 * J
    draws, ng, err = np.zeros((Cn, nw + ns, 7 + D)), np.zeros(Cn, dtype=np.int64), np.zeros(Cn, dtype=np.int32)
    launches = lib.irt_simt_run(Cn, I, J, vp(y), C.c_uint(seed), nw, ns, max_depth, C.c_longlong(tail_after), vp(draws), vp(ng), vp(err))
    return draws, ng, err, launches


@pytest.mark.parametrize("shape", [(2, 20, 6, 8, 4), (1, 290, 5, 6, 3)], ids=["D36", "D304"])
def test_latency_kernel_finishes_chains_like_the_lock_step_engine(lib, shape):
    Cn, I, J, nw, ns = shape   # D = 36: a fraction of the CTA's threads hold an element; D = 304: some threads hold two
    base = _run(lib, Cn, I, J, nw, ns, 4, -1)
    assert np.all(base[2] == 0) and np.all(base[1] > 8) and np.all(np.isfinite(base[0]))
    assert len(np.unique(base[0][:, :, 7])) >= 3      # the chains moved
    for tail_after in (0, 7):
        t = _run(lib, Cn, I, J, nw, ns, 4, tail_after)
        assert np.array_equal(t[0], base[0]) and np.array_equal(t[1], base[1]) and np.all(t[2] == 0)
        assert t[3] == 2 * tail_after + 3   # tail_after x (advance + gradient), advance, latency kernel, the advance that finds every chain idle


def test_cta_per_chain_advance_equals_warp_per_chain_advance(lib):
    """advance_cta.cu (long parameter vectors: the CTA's 256 threads share the element-wise work of the leaf path) against
    advance_kernel (one warp per chain): identical draws and gradient counts."""
    Cn, I, J, nw, ns = 2, 20, 6, 8, 4
    lib.irt_simt_set_advance_cta(0)
    base = _run(lib, Cn, I, J, nw, ns, 4, -1)
    lib.irt_simt_set_advance_cta(1)
    try:
        t = _run(lib, Cn, I, J, nw, ns, 4, -1)
    finally:
        lib.irt_simt_set_advance_cta(0)
    assert np.array_equal(t[0], base[0]) and np.array_equal(t[1], base[1]) and np.all(t[2] == 0) and t[3] == base[3]


def test_straggler_kernel_has_no_race():
    exe = os.path.join(EMUL, "irt_simt_tsan.exe")
    if _stale(exe):
        r = subprocess.run(CXX + ["-fsanitize=thread", "-DIRT_SIMT_MAIN", SRC, "-o", exe], capture_output=True, text=True)
        if r.returncode != 0:
            pytest.skip("no ThreadSanitizer tool chain: " + r.stderr[-300:])
    env = dict(os.environ, TSAN_OPTIONS="exitcode=0 halt_on_error=0")
    r = subprocess.run([exe, "2", "12", "5", "4", "2", "3"], capture_output=True, text=True, env=env, timeout=900, cwd=ROOT)
    if "FATAL: ThreadSanitizer" in r.stderr and "data race" not in r.stderr:
        pytest.skip("ThreadSanitizer cannot run here: " + r.stderr[-300:])
    assert r.returncode == 0 and "agree" in r.stdout and "\nsimt: " not in "\n" + r.stderr, r.stderr[-2000:]
    # the only accepted reports: the 32 same-value stores of the chain's potential (see tests/test_chain_kernels_simt.py)
    from test_chain_kernels_simt import SAME_VALUE_STORE, _top_frames
    for rep in r.stderr.split("WARNING: ThreadSanitizer: data race")[1:]:
        frames = _top_frames(rep)
        assert len(frames) >= 2 and all(SAME_VALUE_STORE.search(f) for f in frames[:2]), rep[:3000]
