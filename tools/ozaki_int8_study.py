#!/usr/bin/env python
"""Numerics study for DESIGN.md 9(2c): the two contractions of the fused GLM gradient (eta = X b, G = X^T d) computed with
INT8 x INT8 -> INT32 products only (what tcgen05.mma kind::i8 does, exactly), operands split into 7-bit slices (Ozaki scheme).

Every product below is an exact integer matrix product (numpy int64 stands in for the INT32 accumulators; the bound on their
magnitude is printed), recombined in fp64.  Question answered: how many slices of X, b and d are needed so that the log-density
and the gradient of the c3 workload stay within the 1e-10 gate -- and with what margin -- and what that costs in INT8 operations
relative to the 2 x 2 N K C fp64 flops of the DMMA kernel.

    python tools/ozaki_int8_study.py            # ~2 min on a CPU; prints a table
"""
import numpy as np

BITS = 7   # magnitude bits per slice: slices are integers in [-64, 64] (two's complement int8 holds them)


def split(A, axis, n_slices):
    """A = scale * sum_s 2^(-BITS (s+1)) S_s + R with integer slices |S_s| <= 2^(BITS-1) and |R| <= scale 2^(-BITS n_slices) / 2.
    `scale` is a power of two >= max |A| along `axis` (one exponent per row / column, kept in fp64)."""
    m = np.max(np.abs(A), axis=axis, keepdims=True)
    scale = 2.0 ** np.ceil(np.log2(np.where(m > 0, m, 1.0)))
    r = A / scale                      # in [-1, 1], exact (power-of-two scaling)
    slices = []
    for _ in range(n_slices):
        r = r * 2.0 ** BITS            # exact
        s = np.rint(r)                 # |s| <= 64 after the first step (|r| <= 1/2 afterwards)
        slices.append(s.astype(np.int64))
        r = r - s                      # exact in fp64
    return scale, slices


def matmul_sliced(As, Bs, sa, sb, max_order):
    """sum over slice pairs (i, j) with i + j <= max_order of 2^(-BITS (i + j + 2)) A_i B_j, scaled; also the largest |int32|"""
    acc = None
    peak = 0
    n = 0
    for order in range(max_order + 1):
        diag = None
        for i in range(len(As)):
            j = order - i
            if 0 <= j < len(Bs):
                p = As[i] @ Bs[j]      # exact integer product
                n += 1
                diag = p if diag is None else diag + p
        if diag is None:
            continue
        peak = max(peak, int(np.max(np.abs(diag))))
        term = diag.astype(np.float64) * 2.0 ** (-BITS * (order + 2))
        acc = term if acc is None else acc + term
    return acc * sa * sb, peak, n


def study(N=20000, K=100, C=64, seed=0):   # ~2 min; N=4000 runs in 25 s with the same picture
    rng = np.random.default_rng(seed)
    X = np.concatenate([np.ones((N, 1)), rng.normal(size=(N, K))], axis=1)        # c3 recipe: X ~ N(0, 1), intercept column
    beta_true = np.concatenate([[0.3], rng.normal(size=K) * 0.3])
    y = (rng.random(N) < 1 / (1 + np.exp(-(X @ beta_true)))).astype(np.float64)
    B = beta_true[:, None] + 0.05 * rng.normal(size=(K + 1, C))                      # chains scattered around the mode
    sgn = 2 * y - 1

    def lp_grad(eta):
        t = sgn[:, None] * eta
        return -np.logaddexp(0, -t).sum(0), sgn[:, None] / (1 + np.exp(t))

    Xl, Bl = X.astype(np.longdouble), B.astype(np.longdouble)
    eta_ref = (Xl @ Bl)
    lp_ref, d_ref = lp_grad(eta_ref.astype(np.float64))
    g_ref = (Xl.T @ d_ref.astype(np.longdouble)).astype(np.float64)
    eta64 = X @ B
    lp64, d64 = lp_grad(eta64)
    g64 = X.T @ d64
    print(f"N={N} K={K} C={C}; fp64 matmul vs 80-bit: lp rel {np.max(np.abs(lp64 - lp_ref) / np.abs(lp_ref)):.1e}, grad rel {np.max(np.abs(g64 - g_ref)) / np.max(np.abs(g_ref)):.1e}")
    print("slices(X,b,d) order | int8 GEMMs A+B | x fp64-flop count | max|int32| A, B | lp rel err | grad rel err (of max |g|)")
    for ns in (3, 4, 5, 6, 7, 8):
        sx, Xs = split(X, 1, ns)                 # one exponent per row of X (static data: sliced once per data set)
        sb, Bs = split(B, 0, ns)                 # one exponent per chain
        order = ns - 1
        eta, peakA, nA = matmul_sliced(Xs, Bs, sx, sb, order)
        lp, d = lp_grad(eta)
        sd, Ds = split(d, 0, ns)                 # derivative tile: one exponent per chain (column), sliced in the epilogue
        XsT = [s.T for s in Xs]
        # G = X^T d = sum_rows sx[row] * Xs[row] * d[row]: the row scale of X is folded into d before slicing
        sd2, Ds2 = split(d * sx, 0, ns)
        G, peakB, nB = matmul_sliced(XsT, Ds2, 1.0, sd2, order)
        e_lp = np.max(np.abs(lp - lp_ref) / np.abs(lp_ref))
        e_g = np.max(np.abs(G - g_ref)) / np.max(np.abs(g_ref))
        print(f"   {ns}            {order}   |   {nA:2d} + {nB:2d}      |      {(nA + nB) / 2:5.1f}         | {peakA:9d} {peakB:11d} |  {e_lp:.1e}  |  {e_g:.1e}")


if __name__ == "__main__":
    study()
