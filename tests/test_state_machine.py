"""CPU suite: the DEVICE state machine (cmdstan_b200/csrc/nuts.cuh: iterative tree, resumable, cached gradients),
compiled for the host with a one-lane warp (tests/host_emul), must reproduce the oracle's recursive formulation bit for
bit -- same draws, same adapted step size and metric -- over warm-up + sampling, for every option that changes the RNG
consumption order (jitter, thinning, save_warmup, max_depth, user inits, fixed metric, adapt off)."""
import numpy as np
import pytest

import cases
import emul_lib as E
from oracle import oracle as O

MODELS = [("bernoulli", 1, 1, cases.BERNOULLI), ("std_normal", 0, 3, {"D": 3}), ("eight_schools_nc", 2, 10, cases.EIGHT_SCHOOLS),
          ("std_normal", 0, 50, {"D": 50})]
OPTS = [
    dict(seed=1234, chain_id=1),
    dict(seed=7, chain_id=3, save_warmup=1, num_warmup=300, num_samples=200, delta=0.95),
    dict(seed=99, chain_id=2, num_warmup=150, num_samples=100, stepsize_jitter=0.3, num_thin=3, save_warmup=1),
    dict(seed=5, chain_id=1, num_warmup=30, num_samples=50, max_depth=4),
    dict(seed=6, chain_id=9, num_warmup=10, num_samples=30),                       # < 20: no variance adaptation
    dict(seed=8, chain_id=1, num_warmup=60, num_samples=30, save_warmup=1),        # 15/75/10 fallback windows
    dict(seed=11, chain_id=4, num_warmup=0, num_samples=40, adapt_engaged=0, stepsize=0.3),
    dict(seed=12, chain_id=2, num_warmup=100, num_samples=20, init_radius=0.0),
    dict(seed=13, chain_id=2, num_warmup=100, num_samples=20, term_buffer=0, init_buffer=50, window=50),
    dict(seed=14, chain_id=3, num_warmup=200, num_samples=50, unit_e=1, save_warmup=1),            # hmc_nuts_unit_e_adapt
    dict(seed=15, chain_id=1, num_warmup=20, num_samples=50, unit_e=1, adapt_engaged=0, stepsize=0.2),  # hmc_nuts_unit_e
    dict(seed=16, chain_id=2, num_warmup=200, num_samples=60, dense_e=1, save_warmup=1),           # hmc_nuts_dense_e_adapt
    dict(seed=17, chain_id=1, num_warmup=60, num_samples=30, dense_e=1, stepsize_jitter=0.2),      # fallback windows, jitter
    dict(seed=18, chain_id=5, num_warmup=10, num_samples=40, dense_e=1, adapt_engaged=0, stepsize=0.3),  # hmc_nuts_dense_e
    dict(seed=19, chain_id=2, num_warmup=150, num_samples=60, save_warmup=1, num_thin=2, diag=1),  # diagnostic rows: q, p, g of every saved draw
    dict(seed=20, chain_id=1, num_warmup=60, num_samples=40, save_warmup=1, diag=1, dense_e=1),
    dict(seed=21, chain_id=3, num_warmup=0, num_samples=50, adapt_engaged=0, stepsize=0.4, max_depth=3, diag=1),
    dict(seed=22, chain_id=1, num_warmup=150, num_samples=60, save_warmup=1, static_hmc=1, int_time=3.5, stepsize_jitter=0.2),   # hmc_static_diag_e_adapt
    dict(seed=23, chain_id=2, num_warmup=20, num_samples=60, static_hmc=1, int_time=1.5, adapt_engaged=0, stepsize=0.3),          # hmc_static_diag_e
    dict(seed=24, chain_id=1, num_warmup=150, num_samples=30, static_hmc=1, int_time=2.0, dense_e=1, save_warmup=1),              # hmc_static_dense_e_adapt
    dict(seed=25, chain_id=4, num_warmup=100, num_samples=30, static_hmc=1, int_time=0.05, unit_e=1, num_thin=2),                  # L = 1; hmc_static_unit_e_adapt
]


@pytest.mark.parametrize("mi", range(len(MODELS)))
@pytest.mark.parametrize("oi", range(len(OPTS)))
def test_device_state_machine_equals_oracle(mi, oi):
    name, kind, D, data = MODELS[mi]
    kw = OPTS[oi]
    a = O.run_chain(O.Model(name, data), **kw)
    b = E.run_chain(kind, D, data, **kw)
    assert a["rc"] == b["rc"] == 0
    assert np.array_equal(a["draws"], b["draws"])
    assert a["stepsize"] == b["stepsize"] and np.array_equal(a["inv_metric"], b["inv_metric"])
    assert b["n_grad"] < a["grad_evals"]  # the device caches (V, g) of the current sample


def test_user_inits_and_metric():
    name, kind, D, data = MODELS[2]
    q0 = np.linspace(-1, 1, D)
    im = np.linspace(0.5, 2.0, D)
    kw = dict(seed=3, chain_id=1, num_warmup=80, num_samples=40, save_warmup=1)
    a = O.run_chain(O.Model(name, data), q0=q0, inv_metric=im, **kw)
    b = E.run_chain(kind, D, data, q0=q0, inv_metric=im, **kw)
    assert np.array_equal(a["draws"], b["draws"])


def test_user_dense_metric():
    """user-supplied dense inverse metric (metric_file of hmc_nuts_dense_e[_adapt]): a correlated SPD matrix."""
    name, kind, D, data = MODELS[2]
    rng = np.random.default_rng(4)
    A = rng.normal(size=(D, D))
    im = A @ A.T / D + np.eye(D)
    for kw in (dict(seed=3, chain_id=1, num_warmup=120, num_samples=40, save_warmup=1, dense_e=1),
               dict(seed=4, chain_id=2, num_warmup=0, num_samples=60, adapt_engaged=0, stepsize=0.15, dense_e=1)):
        a = O.run_chain(O.Model(name, data), inv_metric=im, **kw)
        b = E.run_chain(kind, D, data, inv_metric=im, **kw)
        assert a["rc"] == b["rc"] == 0
        assert np.array_equal(a["draws"], b["draws"]) and np.array_equal(a["inv_metric"], b["inv_metric"])
    # not positive definite: refused (error_codes::CONFIG) by both
    bad = np.eye(D)
    bad[0, 0] = -1.0
    assert O.run_chain(O.Model(name, data), inv_metric=bad, seed=1, num_warmup=20, num_samples=5, dense_e=1)["rc"] == 78
    assert E.run_chain(kind, D, data, inv_metric=bad, seed=1, num_warmup=20, num_samples=5, dense_e=1)["rc"] == 78


def _nan_code(code):
    return np.frombuffer(np.uint64(0x7ff8000000000000 | code).tobytes(), dtype=np.float64)[0]


def test_partial_user_inits():
    """NaN-coded "not given" entries of the user's initial vector (partial inits): the device and the oracle draw the random
    context for every parameter, keep the user's values, and round-trip the others through their transform (code 1: lower
    bound 0, code 2: unit interval); with a partial vector initialisation may retry like a random one."""
    name, kind, D, data = MODELS[2]   # eight schools: mu free, tau lower-bounded, theta_tilde free
    q0 = np.array([0.3, _nan_code(1)] + [_nan_code(0)] * 4 + [0.1, -0.2, 0.3, _nan_code(0)])
    for kw in (dict(seed=3, chain_id=1, num_warmup=60, num_samples=20, save_warmup=1), dict(seed=4, chain_id=2, num_warmup=30, num_samples=10, init_radius=0.0),
               dict(seed=5, chain_id=1, num_warmup=40, num_samples=10, init_radius=0.5, dense_e=1)):
        a = O.run_chain(O.Model(name, data), q0=q0, **kw)
        b = E.run_chain(kind, D, data, q0=q0, **kw)
        assert a["rc"] == b["rc"] == 0
        assert np.array_equal(a["draws"], b["draws"]) and a["stepsize"] == b["stepsize"]
    name, kind, D, data = MODELS[0]   # bernoulli: theta in (0, 1)
    q0 = np.array([_nan_code(2)])
    a = O.run_chain(O.Model(name, data), q0=q0, seed=9, num_warmup=30, num_samples=10)
    b = E.run_chain(kind, D, data, q0=q0, seed=9, num_warmup=30, num_samples=10)
    assert a["rc"] == b["rc"] == 0 and np.array_equal(a["draws"], b["draws"])


def test_init_failure_code():
    """eight schools with sigma so small that every U(-2,2) init overflows is not constructible here; instead check the
    documented failure path with an impossible user init: lp = -inf => error_codes::CONFIG (78)."""
    name, kind, D, data = MODELS[0]
    a = O.run_chain(O.Model(name, data), q0=[np.inf], seed=1, num_warmup=20, num_samples=5)
    b = E.run_chain(kind, D, data, q0=[np.inf], seed=1, num_warmup=20, num_samples=5)
    assert a["rc"] == 78 and b["rc"] == 78


def _random_config(rng):
    nw = int(rng.choice([0, 5, 19, 20, 37, 60, 150, 151, 260]))
    adapt = int(nw > 0 and rng.uniform() < 0.8)
    kw = dict(seed=int(rng.integers(0, 2 ** 31)), chain_id=int(rng.integers(0, 50)), num_warmup=nw, num_samples=int(rng.integers(1, 60)),
              num_thin=int(rng.choice([1, 1, 2, 5])), save_warmup=int(rng.uniform() < 0.5), stepsize=float(rng.choice([1.0, 0.1, 2.5])),
              stepsize_jitter=float(rng.choice([0.0, 0.0, 0.4])), max_depth=int(rng.choice([1, 3, 6, 10])), delta=float(rng.choice([0.6, 0.8, 0.95])),
              gamma=float(rng.choice([0.05, 0.1])), kappa=float(rng.choice([0.75, 0.6])), t0=float(rng.choice([10.0, 3.0])),
              init_buffer=int(rng.choice([75, 10, 0])), term_buffer=int(rng.choice([50, 5, 0])), window=int(rng.choice([25, 7])),
              init_radius=float(rng.choice([2.0, 0.5, 0.0])), adapt_engaged=adapt)
    metric = rng.choice(["diag", "unit", "dense"])
    if metric == "unit":
        kw["unit_e"] = 1
    elif metric == "dense":
        kw["dense_e"] = 1
    if rng.uniform() < 0.3:
        kw["diag"] = 1
    elif rng.uniform() < 0.3:
        kw["static_hmc"] = 1
        kw["int_time"] = float(rng.choice([0.3, 2.0, 6.28]))
    return kw


@pytest.mark.parametrize("case", range(36))
def test_device_state_machine_equals_oracle_random_configs(case):
    """randomised option matrix (every argument of the service that changes control flow or RNG consumption: window
    parameters incl. degenerate ones, thinning, jitter, depth caps, all three metrics, adaptation on/off): the device
    state machine and the oracle must agree bit for bit on every draw, the adapted step size and the metric."""
    rng = np.random.default_rng(1000 + case)
    name, kind, D, data = MODELS[case % len(MODELS)]
    kw = _random_config(rng)
    a = O.run_chain(O.Model(name, data), **kw)
    b = E.run_chain(kind, D, data, **kw)
    assert a["rc"] == b["rc"]
    if a["rc"] == 0:
        assert np.array_equal(a["draws"], b["draws"], equal_nan=True)
        assert (a["stepsize"] == b["stepsize"] or (np.isnan(a["stepsize"]) and np.isnan(b["stepsize"])))
        assert np.array_equal(a["inv_metric"], b["inv_metric"], equal_nan=True)
