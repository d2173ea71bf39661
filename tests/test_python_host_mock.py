"""CPU suite: the Python host (cmdstan_b200.Engine and the service mirrors hmc_nuts_*: ctypes struct layout, data plumbing by
dtype, inits / metric shapes, fetch windows, error codes) driven over the TEST-ONLY mock of the C ABI that serves it from the CPU
oracle (tests/host_emul/mock_b200nuts.cpp).  Everything must then equal the oracle called directly.  The product library is
swapped for the mock by monkeypatching inside the test only; the package itself has no such switch."""
import os
import subprocess

import numpy as np
import pytest

import cases
from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture()
def B(monkeypatch):
    import cmdstan_b200
    from cmdstan_b200 import _build
    O.lib()
    mock_dir = os.path.join(HERE, "host_emul", "mock")
    os.makedirs(mock_dir, exist_ok=True)
    so = os.path.join(mock_dir, "libb200nuts.so")
    src = os.path.join(HERE, "host_emul", "mock_b200nuts.cpp")
    deps = [src, os.path.join(ROOT, "include", "b200nuts.h"), os.path.join(ROOT, "oracle", "nuts_oracle.h"), os.path.join(ROOT, "oracle", "liboracle.so")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-fPIC", "-ffp-contract=off", "-shared", "-o", so, src, "-L" + os.path.join(ROOT, "oracle"),
                               "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    monkeypatch.setattr(_build, "LIB", so)
    monkeypatch.setattr(_build, "needs_build", lambda: False)
    monkeypatch.setattr(cmdstan_b200, "_lib", None)
    yield cmdstan_b200
    cmdstan_b200._lib = None   # the next user loads the real library again


DATA = {"bernoulli": cases.BERNOULLI, "eight_schools_nc": cases.EIGHT_SCHOOLS, "std_normal": {"D": 5}, "logistic": cases.logistic_data(60, 3),
        "poisson_glm": cases.poisson_data(50, 2), "irt_2pl": cases.irt_data(12, 5)}


@pytest.mark.parametrize("name", list(DATA))
def test_service_mirror_equals_oracle(B, name):
    data = DATA[name]
    kw = dict(num_chains=3, random_seed=17, init_chain_id=4, num_warmup=60, num_samples=25, num_thin=2, save_warmup=True, delta=0.9, max_depth=6)
    rc, res = B.hmc_nuts_diag_e_adapt(name, data, **kw)
    assert rc == 0
    m = O.Model(name, data)
    assert res["draws"].shape == (3, 30 + 13, 7 + m.num_outputs()) and len(res["columns"]) == 7 + m.num_outputs()
    for c in range(3):
        r = O.run_chain(m, seed=17, chain_id=4 + c, num_warmup=60, num_samples=25, num_thin=2, save_warmup=1, delta=0.9, max_depth=6)
        want = np.concatenate([r["draws"][:, :7], m.write_array(r["draws"][:, 7:])], axis=1)
        assert np.array_equal(res["draws"][c], want)
        assert res["stepsize"][c] == r["stepsize"] and np.array_equal(res["inv_metric"][c], r["inv_metric"])


def test_sibling_mirrors_inits_metrics_and_windows(B):
    name, data = "eight_schools_nc", cases.EIGHT_SCHOOLS
    m = O.Model(name, data)
    rng = np.random.default_rng(0)
    q0 = rng.uniform(-1, 1, size=(2, 10))
    im = rng.uniform(0.5, 2, size=(2, 10))
    rc, res = B.hmc_nuts_diag_e(name, data, num_chains=2, init=q0, init_inv_metric=im, random_seed=5, num_warmup=10, num_samples=20, stepsize=0.2)
    assert rc == 0 and (res["draws"][:, :, 2] == 0.2).all() and np.array_equal(res["inv_metric"], im)
    for c in range(2):
        r = O.run_chain(m, q0=q0[c], inv_metric=im[c], seed=5, chain_id=1 + c, num_warmup=10, num_samples=20, adapt_engaged=0, stepsize=0.2)
        assert np.array_equal(res["draws"][c][:, 7:], m.write_array(r["draws"][:, 7:]))
    A = rng.normal(size=(10, 10))
    dm = np.broadcast_to(A @ A.T / 10 + np.eye(10), (2, 10, 10)).copy()
    rc, res = B.hmc_nuts_dense_e_adapt(name, data, num_chains=2, init_inv_metric=dm, random_seed=6, num_warmup=150, num_samples=10)
    assert rc == 0 and res["inv_metric"].shape == (2, 10, 10)
    r = O.run_chain(m, inv_metric=dm[1], seed=6, chain_id=2, num_warmup=150, num_samples=10, dense_e=1)
    assert np.array_equal(res["inv_metric"][1], r["inv_metric"]) and np.array_equal(res["draws"][1][:, 7:], m.write_array(r["draws"][:, 7:]))
    rc, res = B.hmc_nuts_unit_e_adapt(name, data, num_chains=1, random_seed=7, num_warmup=40, num_samples=5)
    assert rc == 0 and (res["inv_metric"] == 1).all()
    # engine-level windows and diagnostics
    e = B.Engine(name, data, num_chains=4, random_seed=9, num_warmup=30, num_samples=12, save_warmup=True, store_diagnostics=True)
    e.init_chains()
    e.run(42)
    full = e.fetch_draws(constrained=False)
    assert np.array_equal(e.fetch_draws(chain0=1, nchains=2, row0=7, nrows=9, constrained=False), full[1:3, 7:16])
    dg = e.fetch_diagnostics()
    assert dg.shape == (4, 42, 37) and np.array_equal(dg[:, :, :17], full)
    lp, g = e.log_prob_grad(full[0, :5, 7:])
    lp0, g0 = m.lp_grad(full[0, :5, 7:])
    assert np.array_equal(lp, lp0) and np.array_equal(g, g0) and np.array_equal(dg[0, :5, 27:], -g0)
    err, it, ng = e.chain_status()
    assert (err == 0).all() and (it == 42).all()
    e.close()


def test_error_codes_surface(B):
    rc, res = B.hmc_nuts_diag_e_adapt("bernoulli", cases.BERNOULLI, num_warmup=0)
    assert rc == B.USAGE and res is None
    rc, res = B.hmc_nuts_diag_e_adapt("bernoulli", cases.BERNOULLI, init=np.array([[np.inf]]), num_warmup=20, num_samples=5)
    assert rc == B.CONFIG and res is None   # Initialization failed.


def test_static_hmc_services(B):
    """hmc_static_{diag,unit,dense}_e[_adapt] mirrors over the mock: columns lp__, accept_stat__, stepsize__, int_time__, energy__
    and draws equal to the oracle's static HMC (which is pinned to the reference's engine=static output)."""
    name, data = "eight_schools_nc", cases.EIGHT_SCHOOLS
    m = O.Model(name, data)
    for fn, okw in ((B.hmc_static_diag_e_adapt, {}), (B.hmc_static_unit_e_adapt, dict(unit_e=1)), (B.hmc_static_dense_e, dict(dense_e=1, adapt_engaged=0)),
                    (B.hmc_static_diag_e, dict(adapt_engaged=0))):
        rc, res = fn(name, data, num_chains=2, random_seed=3, num_warmup=60, num_samples=20, save_warmup=True, int_time=1.7, stepsize=0.3)
        assert rc == 0 and res["columns"][:5] == B.STATIC_SAMPLER_COLUMNS and res["columns"][5] == "mu"
        r = O.run_chain(m, seed=3, chain_id=2, num_warmup=60, num_samples=20, save_warmup=1, static_hmc=1, int_time=1.7, stepsize=0.3, **okw)
        want = np.concatenate([r["draws"][:, :3], np.full((80, 1), 1.7), r["draws"][:, 6:7], m.write_array(r["draws"][:, 7:])], axis=1)
        assert np.array_equal(res["draws"][1], want)
