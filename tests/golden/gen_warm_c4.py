"""Warm-start fixture of the reference arm for workload c4 (bench.py --impl reference --workload c4): adapted positions,
inverse metric and step size of 16 chains of the IRT 2PL model at full size, produced on the CPU by the oracle's adaptive NUTS
(the same path the engine runs; on a GPU box `bench.py --dump-warm` produces the c3 fixture from the engine itself).
   python tests/golden/gen_warm_c4.py        (about a minute on 8 cores)"""
import os
import sys
from concurrent.futures import ProcessPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))


def one(chain):
    import cases
    from oracle import oracle as O
    m = O.Model("irt_2pl", cases.irt_data(1000, 100))
    r = O.run_chain(m, seed=20241017, chain_id=1 + chain, num_warmup=150, num_samples=1)
    assert r["rc"] == 0
    return r["draws"][-1, 7:], r["inv_metric"], r["stepsize"]


if __name__ == "__main__":
    with ProcessPoolExecutor(max_workers=os.cpu_count()) as ex:
        res = list(ex.map(one, range(16)))
    np.savez_compressed(os.path.join(HERE, "warm_c4.npz"), q=np.stack([r[0] for r in res]), inv_metric=np.stack([r[1] for r in res]),
                        stepsize=np.array([r[2] for r in res]))
    print("stepsizes", np.round([r[2] for r in res], 4))
