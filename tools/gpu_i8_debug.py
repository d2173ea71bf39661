import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import cases
from oracle import oracle as O
os.environ["B200NUTS_GLM_VARIANT"] = "7"
import cmdstan_b200 as B
for N, K, nq in [(777, 101, 130), (777, 101, 32), (777, 101, 64), (1234, 100, 70), (2000, 100, 20), (30011, 37, 20)]:
    data = cases.logistic_data(N, K)
    m = O.Model("logistic", data)
    q = np.random.default_rng(5).uniform(-1, 1, size=(nq, m.D))
    lp0, g0 = m.lp_grad(q)
    for rep in range(2):
        e = B.Engine("logistic", data)
        lp, g = e.log_prob_grad(q)
        e.close()
        el = np.abs(lp - lp0) / np.abs(lp0)
        eg = np.max(np.abs(g - g0), axis=1) / np.max(np.abs(g0))
        bad = np.where((el > 1e-10) | (eg > 1e-10))[0]
        print(N, K, nq, "rep", rep, "bad chains:", bad.tolist()[:40], "lp err of bad:", np.round(el[bad][:6], 3).tolist(), "grad err:", np.round(eg[bad][:6], 3).tolist(), flush=True)
        if len(bad):
            c = bad[0]
            print("   chain", c, "lp", lp[c], "ref", lp0[c], "ratio", lp[c] / lp0[c], "g[:4]", g[c, :4], "ref", g0[c, :4])
