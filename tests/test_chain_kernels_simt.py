"""CPU suite: the engine's chain kernels (cmdstan_b200/csrc/chains.cuh -- fused_chain_kernel with the chain state in global or
shared memory, advance_kernel + gradient kernel of the lock-step engine) executed on host threads with REAL 32-lane warps by
the SIMT stand-in under tests/host_emul/simt, built like dense.cu (dense metric, diagnostics, static HMC compiled in).

tests/test_state_machine.py drives the same state machine with a one-lane warp and pins it bit for bit to the oracle; this
file covers what a one-lane warp cannot: lane-strided vector loops (D = 50: two elements per lane), the shuffle reductions,
lane-0 random draws broadcast to the warp, the shared-memory staging of the chain state, the compacted request list, and --
with ThreadSanitizer -- the owner-lane rule (an element of a per-chain vector is only touched by its owner lane between two
warp synchronisations).

 * D <= 2: a butterfly sum over at most two non-zero lanes rounds like the sequential sum, so the draws must equal the one-lane
   emulation's (= the oracle's) bit for bit, in every engine;
 * D = 10, 50: the three engines must agree with each other bit for bit (same arithmetic, different staging), a chain must not
   depend on how many chains share the launch, and the first transitions must agree with the oracle to rounding."""
import os
import re
import subprocess

import numpy as np
import pytest

import cases
import emul_lib as E
import simt_chains_lib as S

BERN = ("bernoulli", 1, 1, cases.BERNOULLI)
NORM2 = ("std_normal", 0, 2, {"D": 2})
SCHOOLS = ("eight_schools_nc", 2, 10, cases.EIGHT_SCHOOLS)
NORM50 = ("std_normal", 0, 50, {"D": 50})

OPTS = {
    "diag_e_adapt": dict(seed=5, num_warmup=25, num_samples=10, save_warmup=1),
    "jitter_thin": dict(seed=6, num_warmup=30, num_samples=20, stepsize_jitter=0.3, num_thin=3, save_warmup=1, max_depth=4),
    "unit_e": dict(seed=7, num_warmup=30, num_samples=15, unit_e=1),
    "dense_e_adapt": dict(seed=8, num_warmup=25, num_samples=10, dense_e=1, save_warmup=1),
    "diagnostics": dict(seed=9, num_warmup=25, num_samples=15, diag=1, save_warmup=1),
    "static_adapt": dict(seed=10, num_warmup=30, num_samples=15, static_hmc=1, int_time=1.5, stepsize_jitter=0.2, save_warmup=1),
    "static_dense_fixed": dict(seed=11, num_warmup=5, num_samples=25, static_hmc=1, int_time=0.9, dense_e=1, adapt_engaged=0, stepsize=0.3),
    "no_adapt": dict(seed=12, num_warmup=0, num_samples=30, adapt_engaged=0, stepsize=0.4, init_radius=0.0),
}


@pytest.mark.parametrize("opt", sorted(OPTS))
@pytest.mark.parametrize("model", [BERN, NORM2], ids=lambda m: m[0])
def test_short_vectors_equal_the_oracle_bit_for_bit(model, opt):
    name, kind, D, data = model
    kw = OPTS[opt]
    n = 5 if (name, opt) == ("bernoulli", "diag_e_adapt") else 2   # five chains: two CTAs, the second one partly empty
    ref = [E.run_chain(kind, D, data, chain_id=3 + c, **kw) for c in range(n)]
    for mode in (0, 1, 2):
        r = S.run_chains(mode, n, kind, D, data, chain_id=3, **kw)
        for c in range(n):
            assert r["rc"][c] == ref[c]["rc"] == 0
            assert np.array_equal(r["draws"][c], ref[c]["draws"]), (mode, c)
            assert r["stepsize"][c] == ref[c]["stepsize"] and np.array_equal(r["inv_metric"][c], ref[c]["inv_metric"])
            assert r["n_grad"][c] == ref[c]["n_grad"]


@pytest.mark.parametrize("opt", ["diag_e_adapt", "dense_e_adapt", "diagnostics", "static_adapt", "jitter_thin"])
@pytest.mark.parametrize("model", [SCHOOLS, NORM50], ids=lambda m: "%s%d" % (m[0], m[2]))
def test_engines_agree_and_follow_the_oracle(model, opt):
    name, kind, D, data = model
    kw = dict(OPTS[opt])
    kw.update(num_warmup=min(kw["num_warmup"], 20), num_samples=8)   # 20 warm-up iterations: one metric window (15 % / 75 % / 10 % fallback)
    if D == 50:
        kw.update(num_warmup=min(kw["num_warmup"], 12), num_samples=6)
        if kw.get("dense_e"):
            pytest.skip("covered at D = 10 (a 50 x 50 metric on host threads takes minutes)")
    r = [S.run_chains(mode, 2, kind, D, data, chain_id=1, **kw) for mode in (0, 1, 2)]
    for other in r[1:]:
        assert np.array_equal(other["draws"], r[0]["draws"]) and np.array_equal(other["stepsize"], r[0]["stepsize"])
        assert np.array_equal(other["inv_metric"], r[0]["inv_metric"]) and np.array_equal(other["n_grad"], r[0]["n_grad"])
    assert r[2]["launches"] > 2 * 10 and r[0]["launches"] == 1
    assert np.all(r[0]["rc"] == 0)
    # a chain does not depend on the launch it shares (hmc_nuts_diag_e_adapt_parallel_match_test.cpp:73-153)
    solo = S.run_chains(1, 1, kind, D, data, chain_id=2, **kw)
    assert np.array_equal(solo["draws"][0], r[0]["draws"][1])
    # the first transitions against the one-lane emulation (= the oracle): equal up to the rounding of the 32-lane sums
    ref = E.run_chain(kind, D, data, chain_id=1, **kw)
    k = 6
    a, b = r[0]["draws"][0][:k], ref["draws"][:k]
    assert np.array_equal(a[:, 3:6], b[:, 3:6])           # treedepth / L, n_leapfrog, divergent
    assert np.allclose(a, b, rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize("case", range(10))
def test_random_option_sets(case):
    """randomised option matrix (metrics, diagnostics, static HMC, degenerate windows, depth caps, thinning, jitter, 1-3 chains,
    D = 1, 2, 10, 37): the three engines agree bit for bit, and for D <= 2 with the oracle.  (A 7-minute run of the same generator,
    278 cases, found no mismatch.)"""
    rng = np.random.default_rng(7700 + case)
    name, kind, D, data = [BERN, NORM2, SCHOOLS, ("std_normal", 0, 37, {"D": 37})][int(rng.integers(0, 4))]
    nw = int(rng.choice([0, 5, 19, 20, 37]))
    kw = dict(seed=int(rng.integers(0, 2 ** 31)), chain_id=int(rng.integers(0, 50)), num_warmup=nw, num_samples=int(rng.integers(1, 12)),
              num_thin=int(rng.choice([1, 1, 2])), save_warmup=int(rng.random() < 0.5), stepsize=float(rng.choice([1.0, 0.1, 2.5])),
              stepsize_jitter=float(rng.choice([0.0, 0.4])), max_depth=int(rng.choice([1, 3, 6])), delta=float(rng.choice([0.6, 0.8, 0.95])),
              init_buffer=int(rng.choice([75, 10, 0])), term_buffer=int(rng.choice([50, 5, 0])), window=int(rng.choice([25, 7])),
              init_radius=float(rng.choice([2.0, 0.5, 0.0])), adapt_engaged=int(nw > 0 and rng.random() < 0.8))
    metric = rng.choice(["diag", "unit", "dense"])
    if metric == "unit":
        kw["unit_e"] = 1
    elif metric == "dense" and D <= 10:
        kw["dense_e"] = 1
    if rng.random() < 0.3:
        kw["diag"] = 1
    elif rng.random() < 0.3:
        kw.update(static_hmc=1, int_time=float(rng.choice([0.3, 2.0])))
    n = int(rng.integers(1, 4))
    r = [S.run_chains(mode, n, kind, D, data, **kw) for mode in (0, 1, 2)]
    assert np.all(r[0]["rc"] == 0)
    for other in r[1:]:
        assert np.array_equal(other["draws"], r[0]["draws"]) and np.array_equal(other["stepsize"], r[0]["stepsize"])
        assert np.array_equal(other["inv_metric"], r[0]["inv_metric"])
    if D <= 2:
        for c in range(n):
            e = E.run_chain(kind, D, data, **dict(kw, chain_id=kw["chain_id"] + c))
            assert np.array_equal(e["draws"], r[0]["draws"][c]) and e["stepsize"] == r[0]["stepsize"][c]


def test_user_and_partial_inits_on_32_lanes():
    """init_from_user (out of line since the last hardware session): user values kept, NaN-coded entries take the round trip of
    the random context's draw; bit-identical across the engines and equal to the oracle on the first rows"""
    name, kind, D, data = SCHOOLS
    nan = lambda code: np.frombuffer(np.uint64(0x7ff8000000000000 | code).tobytes(), dtype=np.float64)[0]
    q0 = np.array([0.3, nan(1)] + [nan(0)] * 4 + [0.1, -0.2, 0.3, nan(0)])
    kw = dict(seed=3, num_warmup=10, num_samples=5, save_warmup=1, q0=q0)
    r = [S.run_chains(mode, 2, kind, D, data, chain_id=1, **kw) for mode in (0, 1, 2)]
    assert np.array_equal(r[0]["draws"], r[1]["draws"]) and np.array_equal(r[0]["draws"], r[2]["draws"])
    ref = E.run_chain(kind, D, data, chain_id=1, **kw)
    assert np.allclose(r[0]["draws"][0][:5], ref["draws"][:5], rtol=1e-9, atol=1e-11)
    full = np.linspace(-1, 1, D)
    r = S.run_chains(1, 2, kind, D, data, chain_id=1, seed=4, num_warmup=15, num_samples=5, save_warmup=1, q0=full, init_radius=0.0)
    ref = E.run_chain(kind, D, data, chain_id=1, seed=4, num_warmup=15, num_samples=5, save_warmup=1, q0=full, init_radius=0.0)
    assert np.allclose(r["draws"][0][:5], ref["draws"][:5], rtol=1e-9, atol=1e-11)


# ---- ThreadSanitizer: the owner-lane rule ---------------------------------------------------------------------------------
# All 32 lanes store the chain's potential (a warp-uniform scalar kept in memory) with the same value in the same instruction;
# to a race detector that is a write-write race, on the device it is one warp store.  These are the only reports accepted.
SAME_VALUE_STORE = re.compile(r"\*\s*[cv]\.zV\s*\)?\s*=")


def _tsan_exe():
    exe = os.path.join(S.EMUL, "chains_simt_tsan.exe")
    if S.stale(exe):
        r = subprocess.run(S.CXX + ["-fsanitize=thread", "-DCHAINS_SIMT_MAIN", S.SRC, "-o", exe], capture_output=True, text=True)
        if r.returncode != 0:
            pytest.skip("no ThreadSanitizer tool chain: " + r.stderr[-300:])
    return exe


def _top_frames(report):
    """source lines of frame #0 of each access of one ThreadSanitizer report"""
    out = []
    for m in re.finditer(r"^\s+#0 .*? (\S+?):(\d+)\s", report, flags=re.M):
        path, line = m.group(1), int(m.group(2))
        if os.path.exists(path):
            out.append(open(path).read().splitlines()[line - 1])
        else:
            out.append(path)
    return out


# mode kind D dense diag static iterations
@pytest.mark.parametrize("case", [(1, 2, 10, 0, 0, 0, 6), (2, 2, 10, 1, 1, 0, 5), (0, 0, 50, 0, 0, 1, 4), (1, 2, 10, 1, 0, 1, 4)], ids=lambda c: "-".join(map(str, c)))
def test_owner_lane_rule(case):
    env = dict(os.environ, TSAN_OPTIONS="exitcode=0 halt_on_error=0")
    r = subprocess.run([_tsan_exe()] + [str(a) for a in case], capture_output=True, text=True, env=env, timeout=900, cwd=S.ROOT)
    if "FATAL: ThreadSanitizer" in r.stderr and "data race" not in r.stderr:
        pytest.skip("ThreadSanitizer cannot run here: " + r.stderr[-300:])
    assert r.returncode == 0 and "\nsimt: " not in "\n" + r.stderr and "err 0 0 0" in r.stdout, r.stderr[-2000:]
    reports = r.stderr.split("WARNING: ThreadSanitizer: data race")[1:]
    for rep in reports:
        frames = _top_frames(rep)
        assert len(frames) >= 2 and all(SAME_VALUE_STORE.search(f) for f in frames[:2]), rep[:3000]
