"""CPU suite: index-level emulation of the warp tiling of the experimental GLM kernel v4 (cmdstan_b200/csrc/glm_dmma.cu,
glm_dmma4_kernel) -- the mma.m8n8k4.f64 fragment layouts, the per-warp phase A (64 rows x 8 chains), the in-register link, and the
SHUFFLE exchange that turns the accumulator fragments into phase-B operand fragments -- against plain numpy.  The kernel has not
run on hardware yet; this pins its index arithmetic (the part a compile cannot check).  The same emulator reproduces the
shared-memory exchange of v3 (validated on the GPU) as a control."""
import numpy as np

TR, DK = 64, 24


def dmma(C, a, b):
    """one warp-wide mma.m8n8k4: lane l = 4*gid + tig holds A[gid][tig] = a[l], B[tig][gid] = b[l], C[gid][2*tig + j] = C[l][j]"""
    A = np.zeros((8, 4))
    B = np.zeros((4, 8))
    for l in range(32):
        gid, tig = l >> 2, l & 3
        A[gid, tig] = a[l]
        B[tig, gid] = b[l]
    P = A @ B
    for l in range(32):
        gid, tig = l >> 2, l & 3
        C[l, 0] += P[gid, 2 * tig]
        C[l, 1] += P[gid, 2 * tig + 1]


def run_warp(Xs, Bs, link, exchange):
    """Xs [TR][DK] row tile, Bs [8][DK] the warp's q rows; returns (Eta [TR][8], G [DK][8]) as the kernel's fragments define them"""
    lanes = np.arange(32)
    gid, tig = lanes >> 2, lanes & 3
    acc = np.zeros((8, 32, 2))
    for k0 in range(0, DK, 4):          # phase A
        bv = Bs[gid, k0 + tig]
        for mi in range(8):
            dmma(acc[mi], Xs[mi * 8 + gid, k0 + tig], bv)
    eta = np.zeros((TR, 8))
    for mi in range(8):
        for l in range(32):
            for j in range(2):
                eta[mi * 8 + gid[l], 2 * tig[l] + j] = acc[mi, l, j]
    acc = link(acc)                      # in place: D[row mi*8 + gid][chain 2*tig + j]
    MB = DK // 8
    gacc = np.zeros((MB, 32, 2))
    if exchange == "smem":               # v3: derivative tile through shared memory
        Dw = np.zeros((TR, 8))
        for mi in range(8):
            for l in range(32):
                Dw[mi * 8 + gid[l], 2 * tig[l]:2 * tig[l] + 2] = acc[mi, l]
    for s in range(TR // 4):             # phase B, rows r0 = 4 s .. 4 s + 3
        if exchange == "smem":
            bv = Dw[4 * s + tig, gid]
        else:                            # v4: fetch D[r0 + tig][gid] from the lane that owns it
            src = (((s & 1) * 4 + tig) << 2) | (gid >> 1)
            v0, v1 = acc[s >> 1, src, 0], acc[s >> 1, src, 1]
            bv = np.where(gid & 1, v1, v0)
        for m in range(MB):
            dmma(gacc[m], Xs[4 * s + tig, m * 8 + gid], bv)
    G = np.zeros((DK, 8))
    for m in range(MB):
        for l in range(32):
            G[m * 8 + gid[l], 2 * tig[l]:2 * tig[l] + 2] = gacc[m, l]
    return eta, G


def test_v4_shuffle_exchange_and_v3_control():
    rng = np.random.default_rng(0)
    Xs = rng.normal(size=(TR, DK))
    Bs = rng.normal(size=(8, DK))
    link = lambda acc: np.tanh(acc)      # any elementwise map stands in for the link's derivative
    want_eta = Xs @ Bs.T
    want_G = Xs.T @ np.tanh(want_eta)
    for exchange in ("smem", "shuffle"):
        eta, G = run_warp(Xs, Bs, link, exchange)
        np.testing.assert_allclose(eta, want_eta, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(G, want_G, rtol=1e-12, atol=1e-12)
