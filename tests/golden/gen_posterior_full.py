"""Gate 3 at BASELINE.json's full sizes: posterior summaries of the UNMODIFIED reference (oracle/_ref/<model>, CmdStan 2.36
compiled from /root/reference) for c3 (logistic N = 100k, D = 100) and c4 (IRT 2PL 1000 x 100), committed as small fixtures
because /root/reference does not exist on the GPU box and a full-size reference run is minutes of CPU.

c4 runs the stock adaptive path (1000 warm-up + 1000 draws, 16 chains).  c3: 16 chains x 1000 sampling transitions, warm-started from an adapted state (positions, per-chain inverse metric,
median step size of tests/golden/warm_<workload>.npz; `adapt engaged=0`), i.e. the stock sampler with its adaptation
replaced by a given metric -- the stationary distribution is the posterior whatever the metric, and the start points are
posterior draws, so no burn-in is discarded.  Stored per column (lp__ and every output column): mean, sd, their MCSEs
(S/analyze/mcmc/mcse.hpp through cmdstan_b200/analyze.py, pinned to stansummary), bulk ESS and split R-hat.

   python tests/golden/gen_posterior_full.py c3      (about 10 minutes on 8 cores)
   python tests/golden/gen_posterior_full.py c4
"""
import json
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

CHAINS, DRAWS = 16, 1000


def main(wl):
    import torch
    import cases
    from oracle import oracle as O
    from cmdstan_b200 import analyze as A
    if wl == "c3":
        model, data = "logistic", cases.logistic_data(100000, 100)
    elif wl == "c4":
        model, data = "irt_2pl", cases.irt_data(1000, 100)
    else:
        raise SystemExit("workload must be c3 or c4")
    td = tempfile.mkdtemp(prefix="b200gold_")
    if wl == "c4":
        # c4: the stock adaptive run (1000 warm-up + 1000 draws, random inits): chains warm-started from a 150-transition adaptation
        # had not all reached the typical set (split R-hat 1.19 on every column), so the hierarchical model gets the full path
        t0 = time.time()
        res, wall, _ = O.ref_sample(model, data, num_chains=CHAINS, seed=20261018, num_warmup=1000, num_samples=DRAWS, sig_figs=12,
                                    num_threads=os.cpu_count(), tmpdir=td)
        eps = float(np.median([r[1][-1, 2] for r in res]))
        return finish(wl, res, wall, eps, t0, torch, A)
    z = np.load(os.path.join(HERE, f"warm_{wl}.npz"))
    for c in range(CHAINS):
        q = z["q"][c % len(z["q"])]
        im = z["inv_metric"][c % len(z["inv_metric"])]
        if model == "irt_2pl":
            I, J = data["y"].shape
            init = {"sigma_theta": float(np.exp(q[0])), "theta": q[1:1 + I].tolist(), "sigma_a": float(np.exp(q[1 + I])),
                    "a": np.exp(q[2 + I:2 + I + J]).tolist(), "mu_b": float(q[2 + I + J]), "sigma_b": float(np.exp(q[3 + I + J])),
                    "b": q[4 + I + J:].tolist()}
        else:
            init = {"alpha": float(q[0]), "beta": q[1:].tolist()}
        json.dump(init, open(os.path.join(td, f"init_{c + 1}.json"), "w"))
        json.dump({"inv_metric": im.tolist()}, open(os.path.join(td, f"metric_{c + 1}.json"), "w"))
    eps = float(np.median(z["stepsize"]))
    t0 = time.time()
    res, wall, _ = O.ref_sample(model, data, num_chains=CHAINS, seed=20261018, num_warmup=0, num_samples=DRAWS, sig_figs=12,
                                num_threads=os.cpu_count(), tmpdir=td, adapt_engaged=False, stepsize=eps,
                                metric_file=os.path.join(td, "metric.json"), init=os.path.join(td, "init.json"))
    return finish(wl, res, wall, eps, t0, torch, A)


def finish(wl, res, wall, eps, t0, torch, A):
    draws = np.stack([r[1] for r in res])
    names = res[0][0]
    cols = [0] + list(range(7, draws.shape[2]))
    t = torch.tensor(draws[:, :, cols])
    K = t.shape[2]
    out = dict(columns=np.array([names[k] for k in cols]),
               mean=t.mean(dim=(0, 1)).numpy(), sd=t.reshape(-1, K).std(dim=0, unbiased=True).numpy(),
               mcse_mean=np.array([A.mcse_mean(t[:, :, k]) for k in range(K)]), mcse_sd=np.array([A.mcse_sd(t[:, :, k]) for k in range(K)]),
               ess_bulk=np.array([A.ess_bulk(t[:, :, k]) for k in range(K)]), rhat=np.array([A.split_rhat(t[:, :, k]) for k in range(K)]),
               chains=CHAINS, draws=DRAWS, stepsize=eps, n_leapfrog_mean=float(draws[:, :, 4].mean()), treedepth_mean=float(draws[:, :, 3].mean()),
               divergent=int(draws[:, :, 5].sum()), wall_s=wall)
    np.savez_compressed(os.path.join(HERE, f"ref_posterior_{wl}.npz"), **out)
    print(wl, "done in", round(time.time() - t0), "s; max rhat", float(np.nanmax(out["rhat"])), "min ess", float(np.nanmin(out["ess_bulk"])),
          "mean leapfrogs", out["n_leapfrog_mean"])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "c3")
