"""GPU tests of the device-model plug-in interface (include/b200nuts_model.h) through the C ABI:
(1) a plug-in with the built-in eight-schools arithmetic reproduces the built-in lock-step engine bit for bit (sampler,
write_array, summary); (2) a model the engine does not have (linear regression with unknown sigma,
examples/plugin_models/example_models.cu) passes gate 1 against its closed form in numpy and recovers the posterior."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def B():
    import cmdstan_b200 as B
    from cmdstan_b200 import _build
    B.load_model_library(_build.build_example_plugins())
    return B


def test_plugin_twin_is_bit_identical_to_builtin(B):
    kw = dict(num_chains=37, random_seed=77, num_warmup=150, num_samples=60, save_warmup=True, engine="lockstep")
    a = B.Engine("eight_schools_nc", cases.EIGHT_SCHOOLS, **kw)
    b = B.Engine("eight_schools_plugin", cases.EIGHT_SCHOOLS, **kw)
    assert (a.D, a.K) == (b.D, b.K) == (10, 18)
    q = np.random.default_rng(3).uniform(-2, 2, size=(50, a.D))
    for jac in (True, False):
        la, ga = a.log_prob_grad(q, jacobian=jac)
        lb, gb = b.log_prob_grad(q, jacobian=jac)
        assert np.array_equal(la, lb) and np.array_equal(ga, gb)
    for e in (a, b):
        e.init_chains()
        e.run(210)
    for constrained in (False, True):
        assert np.array_equal(a.fetch_draws(constrained=constrained), b.fetch_draws(constrained=constrained))
    # ragged fetch window (rows of a chain contiguous, chains not): per-chain write_array calls
    assert np.array_equal(a.fetch_draws(chain0=5, nchains=11, row0=150, nrows=17), b.fetch_draws(chain0=5, nchains=11, row0=150, nrows=17))
    sa, sb = a.summary(row0=150, nrows=60), b.summary(row0=150, nrows=60)
    assert np.array_equal(sa, sb, equal_nan=True)
    assert np.array_equal(a.adaptation()[0], b.adaptation()[0])
    assert b.output_names()[:3] == ["v1", "v2", "v3"] and a.output_names()[:3] == ["mu", "tau", "theta_tilde.1"]
    a.close()
    b.close()


def _linreg_data(N=800, K=5, seed=11):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(N, K))
    beta = rng.normal(size=K)
    y = 0.7 + X @ beta + 0.5 * rng.normal(size=N)
    return {"X": X, "y": y}, np.concatenate([[0.7], beta, [0.5]])


def _linreg_lp_grad(data, q, jacobian=True):
    X, y = data["X"], data["y"]
    N, K = X.shape
    lp, g = np.empty(len(q)), np.empty_like(q)
    for s, u in enumerate(q):
        alpha, beta, ls = u[0], u[1:1 + K], u[K + 1]
        sigma = np.exp(ls)
        r = y - alpha - X @ beta
        lp[s] = -alpha ** 2 / 50 - beta @ beta / 8 - sigma ** 2 / 2 - N * ls - r @ r / (2 * sigma ** 2) + (ls if jacobian else 0.0)
        g[s, 0] = -alpha / 25 + r.sum() / sigma ** 2
        g[s, 1:1 + K] = -beta / 4 + X.T @ r / sigma ** 2
        g[s, K + 1] = -sigma ** 2 - N + r @ r / sigma ** 2 + (1.0 if jacobian else 0.0)
    return lp, g


def test_plugin_linreg_gate1_and_posterior(B):
    data, truth = _linreg_data()
    e = B.Engine("linreg", data, num_chains=64, random_seed=5, num_warmup=300, num_samples=200)
    assert e.output_names() == ["alpha"] + [f"beta.{k}" for k in range(1, 6)] + ["sigma"]
    q = np.random.default_rng(2).uniform(-1, 1, size=(40, e.D))
    for jac in (True, False):
        lp, g = e.log_prob_grad(q, jacobian=jac)
        lp0, g0 = _linreg_lp_grad(data, q, jac)
        # gate 1 tolerance (north_star): 1e-10 relative, abs floor 1e-12 / 1e-9
        assert np.max(np.abs(lp - lp0) / np.maximum(1e-12, np.abs(lp0))) < 1e-10
        assert np.max(np.abs(g - g0) / np.maximum(1e-9, np.abs(g0))) < 1e-10
    e.init_chains()
    e.run(500)
    err, it, _ = e.chain_status()
    assert (err == 0).all() and (it == 500).all()
    d = e.fetch_draws()
    assert d.shape == (64, 200, 7 + 7)
    assert d[:, :, 5].mean() < 0.01                       # divergences
    assert np.all(d[:, :, -1] > 0)                         # sigma = exp(log sigma) through the model's write_array
    tab = e.summary()
    mean, sd, rhat = tab[7:, 0], tab[7:, 2], tab[7:, 9]
    assert np.all(rhat < 1.01)
    assert np.all(np.abs(mean - truth) < 5 * sd + 0.02)    # posterior concentrates on the generating values
    # batch independence through the plug-in path: chain 3 alone (stream id 4) == chain 3 of the batch
    e1 = B.Engine("linreg", data, num_chains=1, random_seed=5, init_chain_id=4, num_warmup=300, num_samples=200)
    e1.init_chains()
    e1.run(500)
    assert np.array_equal(e1.fetch_draws()[0], d[3])
    e1.close()
    e.close()


def test_plugin_data_errors(B):
    with pytest.raises(B.B200NutsError) as ei:
        B.Engine("linreg", {"X": np.zeros((4, 2))})
    assert ei.value.code == B.NOINPUT and "y" in str(ei.value)
    with pytest.raises(B.B200NutsError) as ei:
        B.Engine("linreg", {"X": np.zeros((4, 2)), "y": np.zeros(5)})
    assert ei.value.code == B.DATAERR
    with pytest.raises(B.B200NutsError) as ei:
        B.Engine("eight_schools_plugin", cases.EIGHT_SCHOOLS, engine="fused")
    assert ei.value.code == B.CONFIG


def test_plugin_transform_inits(B):
    """vtable.transform_inits (model_base::transform_inits, model_base.hpp:348): user inits on the constrained scale for a plug-in
    model.  The linear regression's sigma has a lower bound: lb_free = log.  A run initialised through it starts where a run
    initialised with the hand-computed unconstrained vector starts (identical draws); elements the user did not give are drawn
    by the engine like the reference's random context; a negative sigma is refused with lb_free's message; a model without the
    callback refuses user inits."""
    data, truth = _linreg_data()
    K = 5
    kw = dict(num_chains=3, random_seed=9, num_warmup=40, num_samples=10, save_warmup=True)
    e = B.Engine("linreg", data, **kw)
    x = np.tile(truth, (3, 1)) + 0.1 * np.arange(3)[:, None]
    q = e.transform_inits(x)
    assert np.array_equal(q[:, :K + 1], x[:, :K + 1]) and np.array_equal(q[:, K + 1], np.log(x[:, K + 1]))
    e.set_inits(q)
    e.init_chains()
    e.run(50)
    d_user = e.fetch_draws(constrained=False)
    e.close()
    e = B.Engine("linreg", data, **kw)
    e.set_inits(np.concatenate([x[:, :K + 1], np.log(x[:, K + 1:])], axis=1))
    e.init_chains()
    e.run(50)
    assert np.array_equal(e.fetch_draws(constrained=False), d_user)
    # partial inits: sigma not given -> marked, the engine draws log sigma ~ U(-2, 2) and keeps the given values
    xp = x.copy()
    xp[:, K + 1] = np.nan
    qp = e.transform_inits(xp)
    assert np.isnan(qp[:, K + 1]).all() and np.array_equal(qp[:, :K + 1], x[:, :K + 1])
    e.close()
    e = B.Engine("linreg", data, num_warmup=0, num_samples=1, adapt_engaged=False, num_chains=3, random_seed=9, stepsize=1e-9)
    e.set_inits(qp)
    e.init_chains()
    e.run(1)
    d0 = e.fetch_draws(constrained=False)[:, 0, 7:]   # a step of 1e-9: the first draw is the initial point
    assert np.allclose(d0[:, :K + 1], x[:, :K + 1], atol=1e-6) and (np.abs(d0[:, K + 1]) <= 2.0).all() and len(set(d0[:, K + 1])) == 3
    xb = x.copy()
    xb[1, K + 1] = -0.3
    with pytest.raises(B.B200NutsError) as ei:
        e.transform_inits(xb)
    assert ei.value.code == B.CONFIG and "lb_free" in str(ei.value)
    e.close()
    s = B.Engine("eight_schools_plugin", cases.EIGHT_SCHOOLS)
    with pytest.raises(B.B200NutsError) as ei:
        s.transform_inits(np.zeros((1, s.K)))
    assert ei.value.code == B.USAGE and "transform_inits" in str(ei.value)
    s.close()
    # built-in models answer the same call with their own transforms
    b = B.Engine("eight_schools_nc", cases.EIGHT_SCHOOLS)
    xs = np.full((1, b.K), np.nan)
    xs[0, :2] = [1.5, 2.0]
    qs = b.transform_inits(xs)
    assert qs[0, 0] == 1.5 and qs[0, 1] == np.log(2.0) and np.isnan(qs[0, 2:]).all()
    b.close()
