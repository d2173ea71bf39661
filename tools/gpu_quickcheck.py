"""Seconds-long hardware check of the code paths changed after the last full GPU suite (no torch import): diagnostics rows,
user / partial inits, the re-organised IRT kernel, and the c2 fused-kernel rate.  Prints one line per check."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import cmdstan_b200 as B  # noqa: E402
from oracle import oracle as O  # noqa: E402


def rel(a, b, floor=1e-6):
    return float(np.max(np.abs(a - b) / np.maximum(floor, np.abs(b))))


def nan_code(code):
    return np.frombuffer(np.uint64(0x7ff8000000000000 | code).tobytes(), dtype=np.float64)[0]


t0 = time.time()
# 1. diagnostics rows, fused engine
e = B.Engine("bernoulli", cases.BERNOULLI, num_chains=2, random_seed=5, num_warmup=40, num_samples=20, save_warmup=True, store_diagnostics=True)
e.init_chains(); e.run(60)
d = e.fetch_diagnostics(); e.close()
r = O.run_chain(O.Model("bernoulli", cases.BERNOULLI), seed=5, chain_id=1, num_warmup=40, num_samples=20, save_warmup=1, diag=1)
print("diag rows bernoulli: shape", d.shape, "tree cols equal", bool(np.array_equal(d[0, :30, 3:6], r["draws"][:30, 3:6])), "rel", rel(d[0, :30], r["draws"][:30]), flush=True)
# 2. user inits (full, partial), lock-step and fused
m = O.Model("eight_schools_nc", cases.EIGHT_SCHOOLS)
for tag, q0 in (("full", np.linspace(-1, 1, 10)), ("partial", np.array([0.3, nan_code(1)] + [nan_code(0)] * 4 + [0.1, -0.2, 0.3, nan_code(0)]))):
    for eng in ("fused", "lockstep"):
        e = B.Engine("eight_schools_nc", cases.EIGHT_SCHOOLS, num_chains=1, random_seed=7, num_warmup=30, num_samples=10, save_warmup=True, engine=eng)
        e.set_inits(q0[None, :]); e.init_chains(); e.run(40)
        d = e.fetch_draws(constrained=False); e.close()
        r = O.run_chain(m, q0=q0, seed=7, chain_id=1, num_warmup=30, num_samples=10, save_warmup=1)
        print(f"user inits {tag} {eng}: tree cols equal", bool(np.array_equal(d[0, :10, 3:6], r["draws"][:10, 3:6])), "rel", rel(d[0, :10], r["draws"][:10]), flush=True)
# 3. IRT gradient after the irt_body.cuh re-organisation
data = cases.irt_data(64, 100)
mi = O.Model("irt_2pl", data)
q = np.random.default_rng(1).uniform(-1.5, 1.5, size=(9, mi.D))
e = B.Engine("irt_2pl", data)
lp, g = e.log_prob_grad(q); e.close()
lp0, g0 = mi.lp_grad(q)
print("irt lp/grad rel", rel(lp, lp0, 1e-12), rel(g, g0, 1e-9), flush=True)
# 4. c2 rate (fused kernel with the interleaved U-turn reductions)
e = B.Engine("eight_schools_nc", cases.EIGHT_SCHOOLS, num_chains=4096, random_seed=20241017, num_warmup=1000, num_samples=205)
e.init_chains(); e.run(1005); e.run(200)
s, _, _ = e.timing(); _, _, ng = e.chain_status(); e.close()
print("c2 fused: %.4f ms per transition of 4096 chains (r01j: 0.1687)" % (1e3 * s / 200), flush=True)
print("total %.1f s" % (time.time() - t0))
