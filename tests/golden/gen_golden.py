#!/usr/bin/env python
"""Generates tests/golden/*.json from the UNMODIFIED reference compiled under oracle/_ref (CmdStan 2.36; see
oracle/build_ref.sh).  Run in the build container (needs /root/reference to have built oracle/_ref):

    python tests/golden/gen_golden.py

Fixtures: (a) log_prob / gradient of every model at seeded random points (method=log_prob jacobian=1 sig_figs=18),
(b) the first rows of sampling CSVs (sig_figs=18, save_warmup=1) for matched-stream checks, (c) a stansummary table.
"""
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
import cases  # noqa: E402
from oracle import oracle as O  # noqa: E402

LPG = {
    "std_normal": {"D": 3},
    "bernoulli": cases.BERNOULLI,
    "eight_schools_nc": cases.EIGHT_SCHOOLS,
    "logistic": cases.logistic_data(200, 7),
    "poisson_glm": cases.poisson_data(150, 5),
    "irt_2pl": cases.irt_data(20, 6),
}
CHAINS = [
    ("bernoulli", cases.BERNOULLI, dict(seed=1234, chain_id=1)),
    ("bernoulli", cases.BERNOULLI, dict(seed=1234, chain_id=2)),
    ("std_normal", {"D": 3}, dict(seed=4839294, chain_id=3)),
    ("eight_schools_nc", cases.EIGHT_SCHOOLS, dict(seed=1234, chain_id=1)),
    ("logistic", LPG["logistic"], dict(seed=77, chain_id=1)),
    ("poisson_glm", LPG["poisson_glm"], dict(seed=78, chain_id=2)),
    ("irt_2pl", LPG["irt_2pl"], dict(seed=79, chain_id=1)),
]


def jsonable(d):
    return {k: np.asarray(v).tolist() for k, v in d.items()}


def main():
    out = {}
    rng = np.random.default_rng(2024)
    for name, data in LPG.items():
        D = O.Model(name, data).D
        q = rng.uniform(-2, 2, size=(12, D))
        lp, g = O.ref_log_prob(name, data, q)
        out[name] = dict(data=jsonable(data), q=q.tolist(), lp=lp.tolist(), grad=g.tolist())
    json.dump(out, open(os.path.join(HERE, "ref_log_prob.json"), "w"))
    chains = []
    for name, data, kw in CHAINS:
        res, _, _ = O.ref_sample(name, data, num_warmup=200, num_samples=100, save_warmup=True, **kw)
        names, draws, stepsize, metric, _ = res[0]
        chains.append(dict(model=name, data=jsonable(data), args=dict(kw, num_warmup=200, num_samples=100), columns=names,
                           rows=draws[:80].tolist(), last_rows=draws[-3:].tolist(), stepsize=stepsize, inv_metric=metric.tolist()))
    json.dump(chains, open(os.path.join(HERE, "ref_chains.json"), "w"))
    # stansummary table (the metric's definition) for a 4-chain bernoulli run
    res, _, files = O.ref_sample("bernoulli", cases.BERNOULLI, num_chains=4, seed=1234, sig_figs=18)
    txt = subprocess.run([os.path.join(O.REF, "stansummary")] + files, capture_output=True, text=True).stdout
    table = {}
    for ln in txt.splitlines():
        p = ln.split()
        if len(p) == 11 and p[0] in ("lp__", "theta", "energy__", "accept_stat__"):
            table[p[0]] = dict(mean=float(p[1]), sd=float(p[3]), ess_bulk=float(p[8]), ess_tail=float(p[9]), rhat=float(p[10]))
    draws = np.stack([r[1] for r in res])
    json.dump(dict(columns=res[0][0], draws=draws[:, :, [0, 1, 6, 7]].tolist(), kept_columns=["lp__", "accept_stat__", "energy__", "theta"],
                   stansummary=table), open(os.path.join(HERE, "ref_stansummary_bernoulli.json"), "w"))
    print("written:", os.listdir(HERE))


if __name__ == "__main__" and "--unit-e" not in sys.argv and "--summary" not in sys.argv and "--dense-e" not in sys.argv:
    main()


def gen_unit_e():
    """metric=unit_e goldens (hmc_nuts_unit_e_adapt / hmc_nuts_unit_e of the compiled reference): first rows at sig_figs=18."""
    out = []
    for name, data, kw in (("bernoulli", cases.BERNOULLI, dict(seed=21, chain_id=1)), ("eight_schools_nc", cases.EIGHT_SCHOOLS, dict(seed=22, chain_id=2))):
        for adapt in (True, False):
            res, _, _ = O.ref_sample(name, data, num_warmup=150, num_samples=50, save_warmup=True, metric="unit_e", adapt_engaged=adapt,
                                     stepsize=None if adapt else 0.25, **kw)
            names, draws, stepsize, metric, _ = res[0]
            out.append(dict(model=name, data=jsonable(data), args=dict(kw, num_warmup=150, num_samples=50, adapt_engaged=int(adapt),
                                                                        stepsize=1.0 if adapt else 0.25), rows=draws[:60].tolist(),
                            last_rows=draws[-3:].tolist(), stepsize=stepsize))
    json.dump(out, open(os.path.join(HERE, "ref_chains_unit_e.json"), "w"))


if __name__ == "__main__" and "--unit-e" in sys.argv:
    gen_unit_e()


def gen_dense_e():
    """metric=dense_e goldens (hmc_nuts_dense_e_adapt / hmc_nuts_dense_e of the compiled reference) at sig_figs=18: every
    saved row + the adapted step size and inverse metric as printed (6 significant digits) in the adaptation block."""
    out = []
    for name, data, kw in (("bernoulli", cases.BERNOULLI, dict(seed=31, chain_id=1)), ("std_normal", {"D": 3}, dict(seed=32, chain_id=2)),
                           ("eight_schools_nc", cases.EIGHT_SCHOOLS, dict(seed=33, chain_id=1))):
        for adapt in (True, False):
            res, _, _ = O.ref_sample(name, data, num_warmup=200, num_samples=100, save_warmup=True, metric="dense_e", adapt_engaged=adapt,
                                     stepsize=None if adapt else 0.25, **kw)
            names, draws, stepsize, metric, _ = res[0]
            out.append(dict(model=name, data=jsonable(data), args=dict(kw, num_warmup=200, num_samples=100, adapt_engaged=int(adapt),
                                                                        stepsize=1.0 if adapt else 0.25), rows=draws.tolist(), stepsize=stepsize,
                            inv_metric=None if metric is None else np.asarray(metric).tolist()))
    json.dump(out, open(os.path.join(HERE, "ref_chains_dense_e.json"), "w"))


if __name__ == "__main__" and "--dense-e" in sys.argv:
    gen_dense_e()


def gen_summary():
    """Full-precision stansummary tables (stansummary -s 12 -c table.csv) of the compiled reference, with the draws they
    were computed from: golden input/output of the posterior-summary restatements (cmdstan_b200/analyze.py and the
    on-device b200nuts_summary).  One case with an even and one with an odd number of draws per chain."""
    import csv
    import tempfile
    out = {}
    for tag, name, data, kw in (("bernoulli_even", "bernoulli", cases.BERNOULLI, dict(num_chains=4, seed=1234, num_warmup=300, num_samples=400)),
                                ("eight_schools_odd", "eight_schools_nc", cases.EIGHT_SCHOOLS, dict(num_chains=5, seed=77, num_warmup=300, num_samples=201))):
        res, _, files = O.ref_sample(name, data, sig_figs=18, **kw)
        td = tempfile.mkdtemp()
        tab = os.path.join(td, "table.csv")
        subprocess.run([os.path.join(O.REF, "stansummary"), "-s", "12", "-c", tab] + files, capture_output=True, text=True, check=True)
        rows = [r for r in csv.reader(l for l in open(tab) if not l.startswith("#"))]
        hdr, body = rows[0], rows[1:]
        out[tag + "_columns"] = np.array(res[0][0])
        out[tag + "_draws"] = np.stack([r[1] for r in res])
        out[tag + "_stat_names"] = np.array(hdr[1:])
        out[tag + "_row_names"] = np.array([r[0] for r in body])
        out[tag + "_table"] = np.array([[float(v) if v not in ("nan", "-nan", "") else np.nan for v in r[1:]] for r in body])
    np.savez_compressed(os.path.join(HERE, "ref_stansummary_full.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__" and "--summary" in sys.argv:
    gen_summary()
