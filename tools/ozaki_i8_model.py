#!/usr/bin/env python
"""Numerical model of csrc/glm_i8.cu (the Ozaki-split GLM gradient on tcgen05 kind::i8), exact in the integers.

Every operand of the two contractions (eta = X q, G = X^T dm) is a fixed-point integer V = rint(v 2^(8 NS - 2)) with |v| <= 1
(39 bits + sign at NS = 5) written in NS balanced base-256 digits (bottom-up: digit = ((V + 128) mod 256) - 128, top digit in
[-64, 64]); the tensor cores form the exact integer products of digit planes and accumulate the pairs (i, j) with i + j = k in
one INT32 accumulator per diagonal k < NS; the diagonals k >= NS (weight 2^-(8 NS) and below) are dropped.  Scales: one power of
two per design column (data, fixed), one per chain for q (max |s q|), none for dm of the logistic link (|dm| <= 1).

    python tools/ozaki_i8_model.py          # the gate-1 metric of the GPU tests (per-component relative error) for NS = 5, 6, 7
"""
import sys

import numpy as np


def digits(V, ns):
    """balanced base-256 digits of the integer array V (int64), most significant first; exact: sum d_i 256^(ns-1-i) == V"""
    out = []
    v = V.copy()
    for _ in range(ns - 1):
        d = ((v + 128) & 255) - 128
        out.append(d)
        v = (v - d) >> 8
    out.append(v)
    assert np.all(np.abs(out[-1]) <= 64)
    return out[::-1]


def col_scale(X):
    m = np.max(np.abs(X), axis=0)
    return 2.0 ** np.ceil(np.log2(np.where(m > 0, m, 1.0)))


def contract(Ad, Bd, ns):
    """sum over i + j < ns of A_i @ B_j * 256^(ns - 1 - (i + j)): the integer S of the kernel (python ints via object would be
    exact beyond 2^63; int64 suffices for the sizes here: |S| < 2^62)"""
    S = None
    for k in range(ns):
        diag = None
        for i in range(k + 1):
            p = Ad[i] @ Bd[k - i]
            diag = p if diag is None else diag + p
        assert np.max(np.abs(diag)) < 2 ** 31, "INT32 accumulator overflow"
        t = diag * (1 << (8 * (ns - 1 - k)))
        S = t if S is None else S + t
    return S


def lp_grad_model(X, y, Q, ns):
    """X [N][D] incl. the intercept column, Q [C][D]; returns the log-likelihood and its gradient as the kernel forms them"""
    F = 8 * ns - 2
    s = col_scale(X)
    Vx = np.rint(X / s * 2.0 ** F).astype(np.int64)
    Xd = digits(Vx, ns)
    qs = Q * s
    eq = np.ceil(np.log2(np.max(np.abs(qs), axis=1, keepdims=True)))
    Vq = np.rint(qs / 2.0 ** eq * 2.0 ** F).astype(np.int64)
    Qd = [d.T for d in digits(Vq, ns)]
    S = contract(Xd, Qd, ns)                                      # [N][C]
    eta = S.astype(np.float64) * 2.0 ** (eq.T - 2 * F + 8 * (ns - 1))
    sgn = (2.0 * y - 1.0)[:, None]
    t = sgn * eta
    ll = -np.logaddexp(0.0, -t)
    dm = sgn / (1.0 + np.exp(t))
    Vd = np.rint(dm * 2.0 ** F).astype(np.int64)
    Dd = digits(Vd, ns)
    Sg = contract([d.T for d in Xd], Dd, ns)                      # [D][C]
    G = (Sg.astype(np.float64) * 2.0 ** (-2 * F + 8 * (ns - 1))) * s[:, None]
    return ll.sum(axis=0), G.T


def lp_grad_exact(X, y, Q):
    Xl, Ql = X.astype(np.longdouble), Q.astype(np.longdouble)
    eta = Xl @ Ql.T
    sgn = (2.0 * y - 1.0).astype(np.longdouble)[:, None]
    t = sgn * eta
    ll = -np.logaddexp(np.longdouble(0), -t)
    dm = sgn / (1 + np.exp(t))
    return ll.sum(axis=0).astype(np.float64), (Xl.T @ dm).T.astype(np.float64)


def rel(a, b, floor):
    return np.max(np.abs(a - b) / np.maximum(floor, np.abs(b)))


def main():
    sys.path.insert(0, "tests")
    import cases
    rows = [(1000, 20, 130, 2.0, 1e-9), (777, 101, 130, 2.0, 1e-9), (5000, 3, 130, 2.0, 1e-9), (1234, 100, 70, 1.0, 1e-9),
            (65, 1, 70, 1.0, 1e-9), (30011, 37, 131, 1.0, 1e-9), (100000, 100, 70, None, 1e-7)]
    print("case (N, K, points)            NS   lp rel      grad rel (per component, floor as in the GPU test)   grad rel to max |g|")
    for N, K, nq, u, floor in rows:
        d = cases.logistic_data(N, K)
        X = np.concatenate([np.ones((N, 1)), d["X"]], axis=1)
        rng = np.random.default_rng(5)
        if u is None:
            Q = np.concatenate([rng.uniform(-2, 2, size=(3, K + 1)), rng.normal(0, 0.1, size=(nq - 3, K + 1))])
        else:
            Q = rng.uniform(-u, u, size=(nq, K + 1))
        lp0, g0 = lp_grad_exact(X, d["y"], Q)
        for ns in (5, 6, 7):
            lp, g = lp_grad_model(X, d["y"], Q, ns)
            print(f"({N:6d}, {K:3d}, {nq:3d})   {ns}   {rel(lp, lp0, 1e-12):.2e}   {rel(g, g0, floor):.2e}   "
                  f"{np.max(np.abs(g - g0)) / np.max(np.abs(g0)):.2e}")


if __name__ == "__main__":
    main()
