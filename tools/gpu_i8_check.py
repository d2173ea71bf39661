#!/usr/bin/env python
"""Hardware check of the INT8 digit-plane GLM kernel (B200NUTS_GLM_VARIANT=7) against the oracle: lp relative error, gradient error
relative to its largest component and per component, for the shapes of the GPU parity tests; then the same points through the FP64
tensor kernel (variant 1) for comparison."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import cases  # noqa: E402
from oracle import oracle as O  # noqa: E402


def rel(a, b, floor):
    return float(np.max(np.abs(a - b) / np.maximum(floor, np.abs(b))))


def main():
    shapes = [(1000, 20, 130), (777, 101, 130), (5000, 3, 130), (1234, 100, 70), (65, 1, 70), (30011, 37, 131), (100000, 100, 70)]
    if len(sys.argv) > 1:
        shapes = shapes[: int(sys.argv[1])]
    for variant in ("7", "1"):
        os.environ["B200NUTS_GLM_VARIANT"] = variant
        import cmdstan_b200 as B
        for N, K, nq in shapes:
            data = cases.logistic_data(N, K)
            m = O.Model("logistic", data)
            rng = np.random.default_rng(5)
            q = rng.uniform(-1, 1, size=(nq, m.D))
            lp0, g0 = m.lp_grad(q)
            e = B.Engine("logistic", data)
            t0 = time.time()
            lp, g = e.log_prob_grad(q)
            dt = time.time() - t0
            e.close()
            print(f"variant {variant} N={N:6d} K={K:3d} points={nq:3d}: lp rel {rel(lp, lp0, 1e-12):.2e}  grad / max|g| {np.max(np.abs(g - g0)) / np.max(np.abs(g0)):.2e}"
                  f"  grad per component (floor 1e-9) {rel(g, g0, 1e-9):.2e}  finite {np.isfinite(lp).all() and np.isfinite(g).all()}  ({dt:.2f} s)", flush=True)


if __name__ == "__main__":
    main()
