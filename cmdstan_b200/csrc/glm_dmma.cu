// FP64 tensor-core (DMMA) kernel of the fused cross-chain GLM gradient (see glm.cu for the algorithm and the
// reference citations).  Both contractions of a row tile run on the tensor cores with mma.sync.m8n8k4.f64
// (tcgen05 has no FP64 kind; DMMA is the FP64 tensor path on sm_100a):
//
//   phase A   Eta[TR x TC]  = Xs[TR x Dk] . Bs[Dk x TC]        (A = X tile, row-major;  B = q of the chain tile)
//   link      lp terms, Dm = dl/deta in registers -> Ds[TR x TC] in shared memory
//   phase B   G[Dk x TC]   += Xs^T[Dk x TR] . Ds[TR x TC]      (A = X tile read transposed from the same smem)
//
// Work decomposition (persistent, one CTA per SM): CTA p owns the fixed row group p (row tiles
// [p T / P, (p+1) T / P), P = min(#SM, T): a function of N only, so a chain's summation order never depends on the batch)
// and loops over the tiles of 64 *requesting* chains (compacted list written by advance_kernel).  Per chain tile the X
// tiles of the group are streamed by the TMA engine (one cp.async.bulk per row, completion on an mbarrier), double
// buffered, the pipeline running across chain-tile boundaries; the X tile is loaded ONCE per (row tile, chain tile) and
// serves both products; the q rows of the chain tile are bulk-copied the same way; Eta / Dm never leave the SM.
//
// Shared-memory leading dimensions are == 4 (mod 8) doubles so that every fragment load (8 rows x 4 k) of a half-warp
// hits 16 distinct 8-byte banks.
//
// The link function is branch-free so that the 16 elements a lane owns are evaluated with full instruction-level
// parallelism (on sm_100 DMMA and scalar FP64 share one pipe: every FP64 instruction saved here is tensor time won):
// exp by range reduction + degree-12 Taylor polynomial, log1p(e) = k ln2 + 2 atanh(z) with one shared reciprocal for z and
// the logistic derivative.  Relative error of each term <= 4e-16 (gate: 1e-10 on lp and gradient vs the reference).
//
// Fragment layout of mma.m8n8k4.f64 (PTX ISA): A a0 = A[lane>>2][lane&3]; B b0 = B[lane&3][lane>>2];
// C c0,c1 = C[lane>>2][2*(lane&3) + {0,1}].
#include "fastmath.cuh"
#include "kernels.h"

namespace b200 {

namespace {

constexpr int TR = 64, TC = 64, NT = 256;
constexpr int DS_LD = TC + 4;

#ifdef B200_SIMT_EMUL  // CPU test-suite only (tests/host_emul/simt): the kernels of this file run on host threads
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) { simt::dmma(d0, d1, a, b); }
#else
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}
#endif

__host__ __device__ inline int xs_ld_for(int Dk) { return Dk + 4; }  // Dk % 8 == 0  =>  ld % 8 == 4

struct GlmTcArgs {
  const double* X;       // [N][ldx], ldx == Dk (multiple of 8), column 0 = 1, zero padded
  const int* y;
  const double* Q;       // [C][ldq]
  const int* list;       // compacted requesting chains, or nullptr (identity)
  const int* n_list;     // device count of requesting chains, or nullptr (n_fixed)
  double* Gp;            // [P][cap][Dk]
  double* LPp;           // [P][cap]
  long long n_row_tiles; // T = ceil(N / TR)
  int N, ldx, Dk, D, ldq, n_fixed, cap, P;
};

// ---- TMA (1-D bulk copy) + mbarrier plumbing ------------------------------------------------------------------------
#ifdef B200_SIMT_EMUL
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) { simt::mbar_init(bar, count); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) { simt::mbar_expect_tx_arrive(bar, bytes); }
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) { simt::mbar_wait(bar, parity); }
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  simt::tma_load_1d(smem_dst, gsrc, bytes, bar);
}
__device__ __forceinline__ void fence_proxy_async() {}
__device__ __forceinline__ void fence_mbarrier_init() {}
#else
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// global -> shared bulk copy (bytes % 16 == 0, both addresses 16-byte aligned), completion on `bar`
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbarrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
#endif

// phase A of one warp: acc[4 row blocks][NB column blocks] += X tile rows . q columns   (NB = column blocks in use)
template <int NB>
__device__ __forceinline__ void phase_a(double (&acc)[4][2][2], const double* xa, const double* ba, int Dk, int xs_ld) {
#pragma unroll 2
  for (int k0 = 0; k0 < Dk; k0 += 4) {
    double av[4], bv[NB];
#pragma unroll
    for (int mi = 0; mi < 4; mi++) av[mi] = xa[mi * 8 * xs_ld + k0];
#pragma unroll
    for (int ni = 0; ni < NB; ni++) bv[ni] = ba[ni * 8 * xs_ld + k0];
#pragma unroll
    for (int mi = 0; mi < 4; mi++)
#pragma unroll
      for (int ni = 0; ni < NB; ni++) dmma(acc[mi][ni][0], acc[mi][ni][1], av[mi], bv[ni]);
  }
}

// phase B of one warp: gacc[MB design-column blocks][NB chain blocks] += X tile^T . Dm tile
template <int NB, int MB>
__device__ __forceinline__ void phase_b(double (&gacc)[MB][2][2], const double* xb, const double* db, int b_mh, int Dk, int xs_ld) {
#pragma unroll 2
  for (int r0 = 0; r0 < TR; r0 += 4) {
    double bv[NB];
#pragma unroll
    for (int ni = 0; ni < NB; ni++) bv[ni] = db[r0 * DS_LD + ni * 8];
#pragma unroll
    for (int m = 0; m < MB; m++) {
      if ((b_mh + m) * 8 < Dk) {  // warp-uniform
        const double av = xb[r0 * xs_ld + m * 8];
#pragma unroll
        for (int ni = 0; ni < NB; ni++) dmma(gacc[m][ni][0], gacc[m][ni][1], av, bv[ni]);
      }
    }
  }
}

// MB: 8-row blocks of G per warp half (Dk <= 16*MB)
template <int LINK, int MB>
__global__ void __launch_bounds__(NT, 1) glm_dmma_kernel(const GlmTcArgs a) {
  B200_DYN_SMEM(smem);
  const int Dk = a.Dk;
  const int xs_ld = xs_ld_for(Dk);
  double* Bs = smem;                               // [TC][xs_ld]     chain-major copy of the q rows of the chain tile
  double* Xs0 = Bs + (size_t)TC * xs_ld;           // [2][TR][xs_ld]  double-buffered X tile
  double* Ds = Xs0 + (size_t)2 * TR * xs_ld;       // [TR][DS_LD]
  double* ys0 = Ds + (size_t)TR * DS_LD;           // [2][TR]  2y-1 / y as double
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(ys0 + 2 * TR);  // fullX[0], fullX[1], fullB
  double* exptab = ys0 + 2 * TR + 4;                // [32]  2^(j/32) (fastmath.cuh)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int p = blockIdx.x;
  const int n_list = a.n_list ? *a.n_list : a.n_fixed;
  const int n_ct = (n_list + TC - 1) / TC;

  // fixed row group of this CTA
  const long long t_begin = (long long)p * a.n_row_tiles / a.P, t_end = (long long)(p + 1) * a.n_row_tiles / a.P;
  const long long row_begin = t_begin * TR;
  long long row_end = t_end * TR;
  if (row_end > a.N) row_end = a.N;
  const int n_rt = (int)(t_end - t_begin);
  const long long n_it = (long long)n_ct * n_rt;
  if (n_it == 0) return;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    fence_mbarrier_init();
  }
  exp_table_load(exptab, tid);
  // columns [ldq, Dk) of the B tile are never written by the bulk copies: zero them once
  for (int e = tid; e < TC * (Dk - a.ldq); e += NT) {
    const int c = e / (Dk - a.ldq), k = a.ldq + e - c * (Dk - a.ldq);
    Bs[c * xs_ld + k] = 0.0;
  }
  fence_proxy_async();
  __syncthreads();

  const unsigned row_bytes = (unsigned)Dk * 8u, q_bytes = (unsigned)a.ldq * 8u;
  // producer side (warp 0): X tile `rt` of this CTA's row group into buffer `buf`
  auto load_x = [&](int rt, int buf) {
    double* Xs = Xs0 + (size_t)buf * TR * xs_ld;
    const long long row0 = row_begin + (long long)rt * TR;
    const int valid = (int)((row_end - row0) < TR ? (row_end - row0) : TR);
    if (lane == 0) mbar_expect_tx(&bars[buf], (unsigned)valid * row_bytes);
    __syncwarp();
#pragma unroll
    for (int h = 0; h < TR / 32; h++) {
      const int r = lane + 32 * h;
      if (r < valid) {
        tma_load_1d(Xs + r * xs_ld, a.X + (size_t)(row0 + r) * a.ldx, row_bytes, &bars[buf]);
      } else {  // rows past the end of the group: X = 0 (no gradient contribution; lp masked in the link)
        for (int k = 0; k < Dk; k += 2) *reinterpret_cast<double2*>(Xs + r * xs_ld + k) = make_double2(0.0, 0.0);
      }
    }
    if (valid < TR) fence_proxy_async();
  };
  auto load_y = [&](int rt, int buf) {  // threads 0..TR-1
    const long long row = row_begin + (long long)rt * TR + tid;
    const int yv = row < row_end ? a.y[row] : 0;
    ys0[buf * TR + tid] = LINK == 0 ? 2.0 * yv - 1.0 : (double)yv;
  };
  auto load_b = [&](int ct) {  // warp 0
    const int nb = n_list - ct * TC < TC ? n_list - ct * TC : TC;
    if (lane == 0) mbar_expect_tx(&bars[2], (unsigned)nb * q_bytes);
    __syncwarp();
#pragma unroll
    for (int h = 0; h < TC / 32; h++) {
      const int c = lane + 32 * h;
      if (c < nb) {
        const int slot = ct * TC + c;
        tma_load_1d(Bs + c * xs_ld, a.Q + (size_t)(a.list ? a.list[slot] : slot) * a.ldq, q_bytes, &bars[2]);
      }
    }
  };

  // phase-A tiling: warp -> 32 rows x 16 chains (4 x 2 mma tiles); phase-B tiling: warp -> MB m-blocks x 16 chains
  const int a_rb = (warp & 1) * 32, a_cb = (warp >> 1) * 16;
  const int b_mh = (warp & 1) * MB, b_cb = (warp >> 1) * 16;

  if (warp == 0) {
    load_b(0);
    load_x(0, 0);
  }
  if (tid < TR) load_y(0, 0);

  double gacc[MB][2][2];
  double lpacc[2][2];       // poisson: running sums of the log-likelihood terms
  LogitAcc lacc[2][2];      // logistic: the same in product form (fastmath.cuh)
  int n_pad = 0;            // logistic: rows past the end of the row group (X == 0 => eta == 0) this lane evaluated
  int ct = 0, rt = 0;
  for (long long it = 0; it < n_it; it++) {
    const int buf = (int)(it & 1);
    if (rt == 0) {
#pragma unroll
      for (int m = 0; m < MB; m++)
#pragma unroll
        for (int n = 0; n < 2; n++) gacc[m][n][0] = gacc[m][n][1] = 0.0;
      lpacc[0][0] = lpacc[0][1] = lpacc[1][0] = lpacc[1][1] = 0.0;
      if (LINK == 0) {
#pragma unroll
        for (int n = 0; n < 2; n++) {
          logit_acc_init(lacc[n][0]);
          logit_acc_init(lacc[n][1]);
        }
        n_pad = 0;
      }
      mbar_wait(&bars[2], (unsigned)(ct & 1));
    }
    mbar_wait(&bars[buf], (unsigned)((it >> 1) & 1));
    __syncthreads();  // every warp is done with the previous iteration: the other X buffer, Ds and ys[buf^1] are free
    if (it + 1 < n_it) {
      const int nrt = rt + 1 < n_rt ? rt + 1 : 0;
      if (warp == 0) load_x(nrt, buf ^ 1);
      if (tid < TR) load_y(nrt, buf ^ 1);
    }
    const double* Xs = Xs0 + (size_t)buf * TR * xs_ld;
    const double* ys = ys0 + buf * TR;
    const long long row0 = row_begin + (long long)rt * TR;
    // 8-chain column blocks of this warp's 16 chains that hold requesting chains (warp-uniform; an underfilled tile --
    // few requesting chains, or the last tile -- skips the tensor and link work of its empty blocks)
    const int in_tile = n_list - ct * TC;
    const int nb = in_tile - a_cb >= 16 ? 2 : (in_tile - a_cb > 8 ? 2 : (in_tile - a_cb > 0 ? 1 : 0));
    // ---- phase A
    double acc[4][2][2];
#pragma unroll
    for (int mi = 0; mi < 4; mi++)
#pragma unroll
      for (int ni = 0; ni < 2; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    {
      const double* xa = Xs + (a_rb + gid) * xs_ld + tig;
      const double* ba = Bs + (a_cb + gid) * xs_ld + tig;
      if (nb == 2) phase_a<2>(acc, xa, ba, Dk, xs_ld);
      else if (nb == 1) phase_a<1>(acc, xa, ba, Dk, xs_ld);
    }
    if (rt == n_rt - 1 && ct + 1 < n_ct) {  // last row tile of this chain tile: the B tile is free after phase A
      __syncthreads();
      if (warp == 0) load_b(ct + 1);
    }
    // ---- link on the accumulator fragments: element (row = a_rb + mi*8 + gid, chain = a_cb + ni*8 + 2*tig + j)
#pragma unroll
    for (int mi = 0; mi < 4; mi++) {
      const int r = a_rb + mi * 8 + gid;
      const bool valid = row0 + r < row_end;
      const double yd = ys[r];
      if (LINK == 0) n_pad += valid ? 0 : 1;
#pragma unroll
      for (int ni = 0; ni < 2; ni++) {
        if (ni >= nb) continue;
        double dv[2];
        if (LINK == 0) {
#pragma unroll
          for (int j = 0; j < 2; j++) dv[j] = logit_term(lacc[ni][j], acc[mi][ni][j], yd, exptab);
        } else {
          double lv[2];
#pragma unroll
          for (int j = 0; j < 2; j++) link_eval<LINK>(acc[mi][ni][j], yd, lv[j], dv[j]);
#pragma unroll
          for (int j = 0; j < 2; j++) lpacc[ni][j] += valid ? lv[j] : 0.0;
        }
        *reinterpret_cast<double2*>(&Ds[r * DS_LD + a_cb + ni * 8 + 2 * tig]) = make_double2(dv[0], dv[1]);
      }
    }
    if (LINK == 0) {  // four more factors in (1, 2] per accumulator: keep the running products' exponents in the integers
#pragma unroll
      for (int ni = 0; ni < 2; ni++) {
        logit_acc_fold(lacc[ni][0]);
        logit_acc_fold(lacc[ni][1]);
      }
    }
    __syncthreads();
    // ---- phase B: G[m][chain] += sum_r Xs[r][m] * Ds[r][chain]
    {
      const double* xb = Xs + tig * xs_ld + b_mh * 8 + gid;
      const double* db = Ds + tig * DS_LD + b_cb + gid;
      if (nb == 2) phase_b<2, MB>(gacc, xb, db, b_mh, Dk, xs_ld);
      else if (nb == 1) phase_b<1, MB>(gacc, xb, db, b_mh, Dk, xs_ld);
    }
    if (rt == n_rt - 1) {
      const int slot0 = ct * TC;
      // ---- partial gradient of this (row group, chain tile)
#pragma unroll
      for (int m = 0; m < MB; m++) {
        const int k = (b_mh + m) * 8 + gid;
        if (k < Dk) {
#pragma unroll
          for (int ni = 0; ni < 2; ni++)
#pragma unroll
            for (int j = 0; j < 2; j++) {
              const int slot = slot0 + b_cb + ni * 8 + 2 * tig + j;
              if (slot < n_list) a.Gp[((size_t)p * a.cap + slot) * Dk + k] = gacc[m][ni][j];
            }
        }
      }
      // ---- partial log-likelihood: sum over the 8 row-lanes (gid) by shuffles, then the two row halves via smem
#pragma unroll
      for (int ni = 0; ni < 2; ni++)
#pragma unroll
        for (int j = 0; j < 2; j++) {
          double v = LINK == 0 ? (ni < nb ? logit_acc_value(lacc[ni][j], n_pad) : 0.0) : lpacc[ni][j];
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          lpacc[ni][j] = v;
        }
      __syncthreads();  // phase B of every warp done: Ds is free
      double* red = Ds;  // [2][TC]
      if (gid == 0) {
#pragma unroll
        for (int ni = 0; ni < 2; ni++)
#pragma unroll
          for (int j = 0; j < 2; j++) red[(warp & 1) * TC + a_cb + ni * 8 + 2 * tig + j] = lpacc[ni][j];
      }
      __syncthreads();
      if (tid < TC && slot0 + tid < n_list) a.LPp[(size_t)p * a.cap + slot0 + tid] = red[tid] + red[TC + tid];
      rt = 0;
      ct++;
    } else {
      rt++;
    }
  }
}


// =====================================================================================================================
// v5 (round 2): v2's register blocking (each warp: 32 rows x 16 chains in phase A, MB design-column blocks x 16 chains in phase
// B -- 6 fragment loads per 8 DMMAs; v3 / v4's 8 x 1 blocking needs 9 and measured 20 % slower on hardware, r01d / r02a), but
// the CTA is TWO GROUPS of four warps that run out of phase.  The r02d capture of v2 shows both tensor phases bound by the DMMA
// pipe (62 % of the time), and the rest spent where all eight warps leave the pipe idle together: the link (13 %), the
// block-wide barriers (11 %), tile set-up.  Here group g owns chain columns [32 g, 32 g + 32) of the chain tile: the two
// groups share the X / q tiles in shared memory but nothing else, synchronise internally with named barriers of 128 threads,
// and group 1 starts half a tile late -- so while one group evaluates the link or waits at its barrier, the other one's DMMAs
// keep the pipe busy (each SM sub-partition holds one warp of each group).  Shared buffers are recycled by the last of the
// eight warps to release them (v3's protocol: hardware-validated in r01d / r02a, race-checked on the CPU rig).
// Summation order per chain is v2's: results are bit-identical to v2.
// =====================================================================================================================
#ifdef B200_SIMT_EMUL
__device__ __forceinline__ void group_bar_sync(int id, int nthreads) { simt::named_bar_sync(id, nthreads); }
__device__ __forceinline__ void group_bar_arrive(int id, int nthreads) { simt::named_bar_arrive(id, nthreads); }
#else
__device__ __forceinline__ void group_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void group_bar_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
#endif

// one arrival per warp on a shared counter; returns true (to the whole warp) in the warp that arrived last of the CTA's eight,
// after resetting the counter.  Orders the warp's earlier shared-memory reads before the refill the last arriver issues.
__device__ __forceinline__ bool warp_arrive_is_last8(int* cnt, int lane) {
  __syncwarp();
  int last = 0;
  if (lane == 0) {
    __threadfence_block();
    last = atomicAdd(cnt, 1) == NT / 32 - 1;
    if (last) {
      *cnt = 0;
      __threadfence_block();
    }
  }
  return __shfl_sync(0xffffffffu, last, 0) != 0;
}

template <int LINK, int MB>
__global__ void __launch_bounds__(NT, 1) glm_dmma5_kernel(const GlmTcArgs a) {
  B200_DYN_SMEM(smem);
  const int Dk = a.Dk;
  const int xs_ld = xs_ld_for(Dk);
  double* Bs = smem;                               // [TC][xs_ld]     chain-major copy of the q rows of the chain tile
  double* Xs0 = Bs + (size_t)TC * xs_ld;           // [2][TR][xs_ld]  double-buffered X tile
  double* Ds = Xs0 + (size_t)2 * TR * xs_ld;       // [TR][DS_LD]     columns [32 g, 32 g + 32): group g's derivative tile
  double* ys0 = Ds + (size_t)TR * DS_LD;           // [2][TR]  2y-1 / y as double
  double* lpred = ys0 + 2 * TR;                    // [2][TC]  log-likelihood partials of the two row halves
  double* exptab = lpred + 2 * TC;                 // [32]     2^(j/32) (fastmath.cuh)
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(exptab + 32);  // fullX[0], fullX[1], fullB
  int* doneX = reinterpret_cast<int*>(bars + 3);   // [2] warps done with X buffer b
  int* doneB = doneX + 2;                          // warps done with the q tile
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int grp = warp >> 2, gw = warp & 3;        // group, warp inside the group
  const int gtid = tid & 127;                      // thread inside the group
  const int p = blockIdx.x;
  const int n_list = a.n_list ? *a.n_list : a.n_fixed;
  const int n_ct = (n_list + TC - 1) / TC;

  // fixed row group of this CTA
  const long long t_begin = (long long)p * a.n_row_tiles / a.P, t_end = (long long)(p + 1) * a.n_row_tiles / a.P;
  const long long row_begin = t_begin * TR;
  long long row_end = t_end * TR;
  if (row_end > a.N) row_end = a.N;
  const int n_rt = (int)(t_end - t_begin);
  const long long n_it = (long long)n_ct * n_rt;
  if (n_it == 0) return;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    doneX[0] = doneX[1] = doneB[0] = 0;
    fence_mbarrier_init();
  }
  exp_table_load(exptab, tid);
  // columns [ldq, Dk) of the B tile are never written by the bulk copies: zero them once
  for (int e = tid; e < TC * (Dk - a.ldq); e += NT) {
    const int c = e / (Dk - a.ldq), k = a.ldq + e - c * (Dk - a.ldq);
    Bs[c * xs_ld + k] = 0.0;
  }
  fence_proxy_async();
  __syncthreads();  // the only block-wide barrier of the kernel

  const unsigned row_bytes = (unsigned)Dk * 8u, q_bytes = (unsigned)a.ldq * 8u;
  // issued by one whole warp: X tile + y slice of iteration `it2` into its buffer
  auto load_x = [&](long long it2) {
    const int buf = (int)(it2 & 1);
    const int rt2 = (int)(it2 % n_rt);
    double* Xs = Xs0 + (size_t)buf * TR * xs_ld;
    const long long row0 = row_begin + (long long)rt2 * TR;
    const int valid = (int)((row_end - row0) < TR ? (row_end - row0) : TR);
    // generic-proxy part of the tile first (y slice, zero rows past the end of the group): released by the arrive below
#pragma unroll
    for (int h = 0; h < TR / 32; h++) {
      const int r = lane + 32 * h;
      const int yv = r < valid ? a.y[row0 + r] : 0;
      ys0[buf * TR + r] = LINK == 0 ? 2.0 * yv - 1.0 : (double)yv;
      if (r >= valid)
        for (int k = 0; k < Dk; k += 2) *reinterpret_cast<double2*>(Xs + r * xs_ld + k) = make_double2(0.0, 0.0);
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_expect_tx(&bars[buf], (unsigned)valid * row_bytes);
    __syncwarp();
#pragma unroll
    for (int h = 0; h < TR / 32; h++) {
      const int r = lane + 32 * h;
      if (r < valid) tma_load_1d(Xs + r * xs_ld, a.X + (size_t)(row0 + r) * a.ldx, row_bytes, &bars[buf]);
    }
  };
  auto load_b = [&](int ct2) {  // one whole warp: q rows of chain tile ct2
    const int nbq = n_list - ct2 * TC < TC ? n_list - ct2 * TC : TC;
    fence_proxy_async();
    if (lane == 0) mbar_expect_tx(&bars[2], (unsigned)nbq * q_bytes);
    __syncwarp();
#pragma unroll
    for (int h = 0; h < TC / 32; h++) {
      const int c = lane + 32 * h;
      if (c < nbq) {
        const int slot = ct2 * TC + c;
        tma_load_1d(Bs + c * xs_ld, a.Q + (size_t)(a.list ? a.list[slot] : slot) * a.ldq, q_bytes, &bars[2]);
      }
    }
  };
  if (warp == 0) {
    load_b(0);
    load_x(0);
    if (n_it > 1) load_x(1);
  }

  // phase-A tiling: warp -> 32 rows x 16 chains (4 x 2 mma tiles); phase-B tiling: warp -> MB m-blocks x 16 chains
  const int a_rb = (gw & 1) * 32, a_cb = grp * 32 + (gw >> 1) * 16;
  const int b_mh = (gw & 1) * MB, b_cb = a_cb;
  const int bar_id = 1 + grp;   // named barrier of this group (barrier 0 is __syncthreads')

  double gacc[MB][2][2];
  double lpacc[2][2];       // poisson: running sums of the log-likelihood terms
  LogitAcc lacc[2][2];      // logistic: the same in product form (fastmath.cuh)
  int n_pad = 0;            // logistic: rows past the end of the row group (X == 0 => eta == 0) this lane evaluated
  int ct = 0, rt = 0;
  for (long long it = 0; it < n_it; it++) {
    const int buf = (int)(it & 1);
    if (rt == 0) {
#pragma unroll
      for (int m = 0; m < MB; m++)
#pragma unroll
        for (int n = 0; n < 2; n++) gacc[m][n][0] = gacc[m][n][1] = 0.0;
      lpacc[0][0] = lpacc[0][1] = lpacc[1][0] = lpacc[1][1] = 0.0;
      if (LINK == 0) {
#pragma unroll
        for (int n = 0; n < 2; n++) {
          logit_acc_init(lacc[n][0]);
          logit_acc_init(lacc[n][1]);
        }
        n_pad = 0;
      }
      mbar_wait(&bars[2], (unsigned)(ct & 1));
    }
    mbar_wait(&bars[buf], (unsigned)((it >> 1) & 1));
    if (it == 0 && grp == 1) group_bar_sync(3, NT);   // group 1 starts when group 0 has finished its first phase A
    const double* Xs = Xs0 + (size_t)buf * TR * xs_ld;
    const double* ys = ys0 + buf * TR;
    const long long row0 = row_begin + (long long)rt * TR;
    // 8-chain column blocks of this warp's 16 chains that hold requesting chains (warp-uniform)
    const int in_tile = n_list - ct * TC;
    const int nb = in_tile - a_cb >= 16 ? 2 : (in_tile - a_cb > 8 ? 2 : (in_tile - a_cb > 0 ? 1 : 0));
    // ---- phase A
    double acc[4][2][2];
#pragma unroll
    for (int mi = 0; mi < 4; mi++)
#pragma unroll
      for (int ni = 0; ni < 2; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    {
      const double* xa = Xs + (a_rb + gid) * xs_ld + tig;
      const double* ba = Bs + (a_cb + gid) * xs_ld + tig;
      if (nb == 2) phase_a<2>(acc, xa, ba, Dk, xs_ld);
      else if (nb == 1) phase_a<1>(acc, xa, ba, Dk, xs_ld);
    }
    if (it == 0 && grp == 0) group_bar_arrive(3, NT);
    if (rt == n_rt - 1 && ct + 1 < n_ct) {  // last row tile of this chain tile: this warp is done with the q tile
      if (warp_arrive_is_last8(doneB, lane)) load_b(ct + 1);
    }
    group_bar_sync(bar_id, 128);  // every warp of the group has finished phase B of the previous tile: Ds (our columns) is free
    // ---- link on the accumulator fragments: element (row = a_rb + mi*8 + gid, chain = a_cb + ni*8 + 2*tig + j)
#pragma unroll
    for (int mi = 0; mi < 4; mi++) {
      const int r = a_rb + mi * 8 + gid;
      const bool valid = row0 + r < row_end;
      const double yd = ys[r];
      if (LINK == 0) n_pad += valid ? 0 : 1;
#pragma unroll
      for (int ni = 0; ni < 2; ni++) {
        if (ni >= nb) continue;
        double dv[2];
        if (LINK == 0) {
#pragma unroll
          for (int j = 0; j < 2; j++) dv[j] = logit_term(lacc[ni][j], acc[mi][ni][j], yd, exptab);
        } else {
          double lv[2];
#pragma unroll
          for (int j = 0; j < 2; j++) link_eval<LINK>(acc[mi][ni][j], yd, lv[j], dv[j]);
#pragma unroll
          for (int j = 0; j < 2; j++) lpacc[ni][j] += valid ? lv[j] : 0.0;
        }
        *reinterpret_cast<double2*>(&Ds[r * DS_LD + a_cb + ni * 8 + 2 * tig]) = make_double2(dv[0], dv[1]);
      }
    }
    if (LINK == 0) {
#pragma unroll
      for (int ni = 0; ni < 2; ni++) {
        logit_acc_fold(lacc[ni][0]);
        logit_acc_fold(lacc[ni][1]);
      }
    }
    group_bar_sync(bar_id, 128);  // the group's derivative tile is complete
    // ---- phase B: G[m][chain] += sum_r Xs[r][m] * Ds[r][chain]
    {
      const double* xb = Xs + tig * xs_ld + b_mh * 8 + gid;
      const double* db = Ds + tig * DS_LD + b_cb + gid;
      if (nb == 2) phase_b<2, MB>(gacc, xb, db, b_mh, Dk, xs_ld);
      else if (nb == 1) phase_b<1, MB>(gacc, xb, db, b_mh, Dk, xs_ld);
    }
    // this warp is done with Xs[buf] and ys[buf]: the last of the eight warps to get here refills them for iteration it + 2
    if (it + 2 < n_it) {
      if (warp_arrive_is_last8(&doneX[buf], lane)) load_x(it + 2);
    } else {
      __syncwarp();
    }
    if (rt == n_rt - 1) {
      const int slot0 = ct * TC;
      // ---- partial gradient of this (row group, chain tile)
#pragma unroll
      for (int m = 0; m < MB; m++) {
        const int k = (b_mh + m) * 8 + gid;
        if (k < Dk) {
#pragma unroll
          for (int ni = 0; ni < 2; ni++)
#pragma unroll
            for (int j = 0; j < 2; j++) {
              const int slot = slot0 + b_cb + ni * 8 + 2 * tig + j;
              if (slot < n_list) a.Gp[((size_t)p * a.cap + slot) * Dk + k] = gacc[m][ni][j];
            }
        }
      }
      // ---- partial log-likelihood: sum over the 8 row-lanes (gid) by shuffles, then the two row halves via shared memory
#pragma unroll
      for (int ni = 0; ni < 2; ni++)
#pragma unroll
        for (int j = 0; j < 2; j++) {
          double v = LINK == 0 ? (ni < nb ? logit_acc_value(lacc[ni][j], n_pad) : 0.0) : lpacc[ni][j];
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          lpacc[ni][j] = v;
        }
      if (gid == 0) {
#pragma unroll
        for (int ni = 0; ni < 2; ni++)
#pragma unroll
          for (int j = 0; j < 2; j++) lpred[(gw & 1) * TC + a_cb + ni * 8 + 2 * tig + j] = lpacc[ni][j];
      }
      group_bar_sync(bar_id, 128);
      if (gtid < 32 && slot0 + grp * 32 + gtid < n_list)
        a.LPp[(size_t)p * a.cap + slot0 + grp * 32 + gtid] = lpred[grp * 32 + gtid] + lpred[TC + grp * 32 + gtid];
      group_bar_sync(bar_id, 128);  // lpred (our columns) is read before the next chain tile's partials are written
      rt = 0;
      ct++;
    } else {
      rt++;
    }
  }
}

size_t glm_tc5_smem_bytes(int Dk) {
  return ((size_t)TC * xs_ld_for(Dk) + (size_t)2 * TR * xs_ld_for(Dk) + (size_t)TR * DS_LD + 2 * TR + 2 * TC + 32 + 3 + 2) * sizeof(double);
}


// =====================================================================================================================
// v6 (round 2): SIXTEEN warps in two out-of-phase groups of eight.  What r02d / r02e established: a warp's DMMA stream carries
// a fixed issue gap after every DMMA.8x8x4 (the NOP ptxas pairs with it; sm_100a has no larger FP64 shape -- ptxas lowers
// m16n8k16 to 8x8x4 too), so the pipe is only saturated while at least TWO warps of an SM sub-partition issue DMMAs.  v2 (eight
// warps) satisfies that inside its tensor phases but idles the pipe whenever its warps do anything else together (38 % of the
// time); v5 (two groups of four) de-phased the groups but left each group with ONE warp per sub-partition, which cannot fill the
// pipe alone.  Here each group has two warps per sub-partition: while one group evaluates the link, waits at its barrier or
// writes partials, the other one keeps the pipe full.  512 threads at <= 128 registers: per-warp tiles shrink to 16 rows x 16
// chains (phase A, 4 fragment loads per 4 DMMAs) and 7 design-column blocks x 8 chains (phase B, 8 loads per 7 DMMAs).
// Row partials are summed over four row quarters (not v2's two halves): same accuracy, not the same bits as v2.
// =====================================================================================================================
constexpr int NT6 = 512, NW6 = NT6 / 32;

__device__ __forceinline__ bool warp_arrive_is_last16(int* cnt, int lane) {
  __syncwarp();
  int last = 0;
  if (lane == 0) {
    __threadfence_block();
    last = atomicAdd(cnt, 1) == NW6 - 1;
    if (last) {
      *cnt = 0;
      __threadfence_block();
    }
  }
  return __shfl_sync(0xffffffffu, last, 0) != 0;
}

template <int LINK, int MB>   // MB: 8-row blocks of G per warp in phase B (Dk <= 16 MB)
__global__ void __launch_bounds__(NT6, 1) glm_dmma6_kernel(const GlmTcArgs a) {
  B200_DYN_SMEM(smem);
  const int Dk = a.Dk;
  const int xs_ld = xs_ld_for(Dk);
  double* Bs = smem;                               // [TC][xs_ld]     chain-major copy of the q rows of the chain tile
  double* Xs0 = Bs + (size_t)TC * xs_ld;           // [2][TR][xs_ld]  double-buffered X tile
  double* Ds = Xs0 + (size_t)2 * TR * xs_ld;       // [TR][DS_LD]     columns [32 g, 32 g + 32): group g's derivative tile
  double* ys0 = Ds + (size_t)TR * DS_LD;           // [2][TR]  2y-1 / y as double
  double* lpred = ys0 + 2 * TR;                    // [4][TC]  log-likelihood partials of the four row quarters
  double* exptab = lpred + 4 * TC;                 // [32]     2^(j/32) (fastmath.cuh)
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(exptab + 32);  // fullX[0], fullX[1], fullB
  int* doneX = reinterpret_cast<int*>(bars + 3);   // [2] warps done with X buffer b
  int* doneB = doneX + 2;                          // warps done with the q tile
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int grp = warp >> 3, gw = warp & 7;        // group, warp inside the group (two warps of each group per sub-partition)
  const int gtid = tid & 255;                      // thread inside the group
  const int p = blockIdx.x;
  const int n_list = a.n_list ? *a.n_list : a.n_fixed;
  const int n_ct = (n_list + TC - 1) / TC;

  const long long t_begin = (long long)p * a.n_row_tiles / a.P, t_end = (long long)(p + 1) * a.n_row_tiles / a.P;
  const long long row_begin = t_begin * TR;
  long long row_end = t_end * TR;
  if (row_end > a.N) row_end = a.N;
  const int n_rt = (int)(t_end - t_begin);
  const long long n_it = (long long)n_ct * n_rt;
  if (n_it == 0) return;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    doneX[0] = doneX[1] = doneB[0] = 0;
    fence_mbarrier_init();
  }
  exp_table_load(exptab, tid);
  for (int e = tid; e < TC * (Dk - a.ldq); e += NT6) {  // columns [ldq, Dk) of the q tile are never written by the bulk copies
    const int c = e / (Dk - a.ldq), k = a.ldq + e - c * (Dk - a.ldq);
    Bs[c * xs_ld + k] = 0.0;
  }
  fence_proxy_async();
  __syncthreads();  // the only block-wide barrier of the kernel

  const unsigned row_bytes = (unsigned)Dk * 8u, q_bytes = (unsigned)a.ldq * 8u;
  auto load_x = [&](long long it2) {  // one whole warp: X tile + y slice of iteration it2 into its buffer
    const int buf = (int)(it2 & 1);
    const int rt2 = (int)(it2 % n_rt);
    double* Xs = Xs0 + (size_t)buf * TR * xs_ld;
    const long long row0 = row_begin + (long long)rt2 * TR;
    const int valid = (int)((row_end - row0) < TR ? (row_end - row0) : TR);
#pragma unroll
    for (int h = 0; h < TR / 32; h++) {
      const int r = lane + 32 * h;
      const int yv = r < valid ? a.y[row0 + r] : 0;
      ys0[buf * TR + r] = LINK == 0 ? 2.0 * yv - 1.0 : (double)yv;
      if (r >= valid)
        for (int k = 0; k < Dk; k += 2) *reinterpret_cast<double2*>(Xs + r * xs_ld + k) = make_double2(0.0, 0.0);
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_expect_tx(&bars[buf], (unsigned)valid * row_bytes);
    __syncwarp();
#pragma unroll
    for (int h = 0; h < TR / 32; h++) {
      const int r = lane + 32 * h;
      if (r < valid) tma_load_1d(Xs + r * xs_ld, a.X + (size_t)(row0 + r) * a.ldx, row_bytes, &bars[buf]);
    }
  };
  auto load_b = [&](int ct2) {  // one whole warp: q rows of chain tile ct2
    const int nbq = n_list - ct2 * TC < TC ? n_list - ct2 * TC : TC;
    fence_proxy_async();
    if (lane == 0) mbar_expect_tx(&bars[2], (unsigned)nbq * q_bytes);
    __syncwarp();
#pragma unroll
    for (int h = 0; h < TC / 32; h++) {
      const int c = lane + 32 * h;
      if (c < nbq) {
        const int slot = ct2 * TC + c;
        tma_load_1d(Bs + c * xs_ld, a.Q + (size_t)(a.list ? a.list[slot] : slot) * a.ldq, q_bytes, &bars[2]);
      }
    }
  };
  if (warp == 0) {
    load_b(0);
    load_x(0);
    if (n_it > 1) load_x(1);
  }

  // phase A: warp -> 16 rows x 16 chains (2 x 2 mma tiles); phase B: warp -> MB design-column blocks x 8 chains
  const int a_rq = gw & 3;                               // row quarter
  const int a_rb = a_rq * 16, a_cb = grp * 32 + (gw >> 2) * 16;
  const int b_mh = (gw >> 2) * MB, b_cb = grp * 32 + (gw & 3) * 8;
  const int bar_id = 1 + grp;

  double gacc[MB][2];
  double lpacc[2][2];
  LogitAcc lacc[2][2];
  int n_pad = 0;
  int ct = 0, rt = 0;
  for (long long it = 0; it < n_it; it++) {
    const int buf = (int)(it & 1);
    if (rt == 0) {
#pragma unroll
      for (int m = 0; m < MB; m++) gacc[m][0] = gacc[m][1] = 0.0;
      lpacc[0][0] = lpacc[0][1] = lpacc[1][0] = lpacc[1][1] = 0.0;
      if (LINK == 0) {
#pragma unroll
        for (int n = 0; n < 2; n++) {
          logit_acc_init(lacc[n][0]);
          logit_acc_init(lacc[n][1]);
        }
        n_pad = 0;
      }
      mbar_wait(&bars[2], (unsigned)(ct & 1));
    }
    mbar_wait(&bars[buf], (unsigned)((it >> 1) & 1));
    if (it == 0 && grp == 1) group_bar_sync(3, NT6);   // group 1 starts when group 0 has finished its first phase A
    const double* Xs = Xs0 + (size_t)buf * TR * xs_ld;
    const double* ys = ys0 + buf * TR;
    const long long row0 = row_begin + (long long)rt * TR;
    const int in_tile = n_list - ct * TC;
    // 8-chain blocks of this warp's 16 phase-A chains that hold requesting chains; does its phase-B block hold any? (warp-uniform)
    const int nb = in_tile - a_cb > 8 ? 2 : (in_tile - a_cb > 0 ? 1 : 0);
    const bool haveB = in_tile - b_cb > 0;
    // ---- phase A: Eta[16 rows][16 chains]
    double acc[2][2][2];
#pragma unroll
    for (int mi = 0; mi < 2; mi++)
#pragma unroll
      for (int ni = 0; ni < 2; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    if (nb > 0) {
      const double* xa = Xs + (a_rb + gid) * xs_ld + tig;
      const double* ba = Bs + (a_cb + gid) * xs_ld + tig;
#pragma unroll 2
      for (int k0 = 0; k0 < Dk; k0 += 4) {
        const double av0 = xa[k0], av1 = xa[8 * xs_ld + k0];
        const double bv0 = ba[k0], bv1 = nb == 2 ? ba[8 * xs_ld + k0] : 0.0;
        dmma(acc[0][0][0], acc[0][0][1], av0, bv0);
        dmma(acc[1][0][0], acc[1][0][1], av1, bv0);
        if (nb == 2) {
          dmma(acc[0][1][0], acc[0][1][1], av0, bv1);
          dmma(acc[1][1][0], acc[1][1][1], av1, bv1);
        }
      }
    }
    if (it == 0 && grp == 0) group_bar_arrive(3, NT6);
    if (rt == n_rt - 1 && ct + 1 < n_ct) {  // last row tile of this chain tile: this warp is done with the q tile
      if (warp_arrive_is_last16(doneB, lane)) load_b(ct + 1);
    }
    group_bar_sync(bar_id, 256);  // every warp of the group has finished phase B of the previous tile: Ds (our columns) is free
    // ---- link: element (row = a_rb + mi*8 + gid, chain = a_cb + ni*8 + 2*tig + j)
#pragma unroll
    for (int mi = 0; mi < 2; mi++) {
      const int r = a_rb + mi * 8 + gid;
      const bool valid = row0 + r < row_end;
      const double yd = ys[r];
      if (LINK == 0) n_pad += valid ? 0 : 1;
#pragma unroll
      for (int ni = 0; ni < 2; ni++) {
        if (ni >= nb) continue;
        double dv[2];
        if (LINK == 0) {
#pragma unroll
          for (int j = 0; j < 2; j++) dv[j] = logit_term(lacc[ni][j], acc[mi][ni][j], yd, exptab);
        } else {
          double lv[2];
#pragma unroll
          for (int j = 0; j < 2; j++) link_eval<LINK>(acc[mi][ni][j], yd, lv[j], dv[j]);
#pragma unroll
          for (int j = 0; j < 2; j++) lpacc[ni][j] += valid ? lv[j] : 0.0;
        }
        *reinterpret_cast<double2*>(&Ds[r * DS_LD + a_cb + ni * 8 + 2 * tig]) = make_double2(dv[0], dv[1]);
      }
    }
    if (LINK == 0) {
#pragma unroll
      for (int ni = 0; ni < 2; ni++) {
        logit_acc_fold(lacc[ni][0]);
        logit_acc_fold(lacc[ni][1]);
      }
    }
    group_bar_sync(bar_id, 256);  // the group's derivative tile is complete
    // ---- phase B: G[m][8 chains] += sum_r Xs[r][m] * Ds[r][chain]
    if (haveB) {
      const double* xb = Xs + tig * xs_ld + b_mh * 8 + gid;
      const double* db = Ds + tig * DS_LD + b_cb + gid;
#pragma unroll 2
      for (int r0 = 0; r0 < TR; r0 += 4) {
        const double bv = db[r0 * DS_LD];
#pragma unroll
        for (int m = 0; m < MB; m++) {
          if ((b_mh + m) * 8 < Dk) {  // warp-uniform
            const double av = xb[r0 * xs_ld + m * 8];
            dmma(gacc[m][0], gacc[m][1], av, bv);
          }
        }
      }
    }
    if (it + 2 < n_it) {
      if (warp_arrive_is_last16(&doneX[buf], lane)) load_x(it + 2);
    } else {
      __syncwarp();
    }
    if (rt == n_rt - 1) {
      const int slot0 = ct * TC;
      if (haveB) {
#pragma unroll
        for (int m = 0; m < MB; m++) {
          const int k = (b_mh + m) * 8 + gid;
          if (k < Dk) {
#pragma unroll
            for (int j = 0; j < 2; j++) {
              const int slot = slot0 + b_cb + 2 * tig + j;
              if (slot < n_list) a.Gp[((size_t)p * a.cap + slot) * Dk + k] = gacc[m][j];
            }
          }
        }
      }
      // ---- partial log-likelihood: sum over the 8 row-lanes (gid) by shuffles, then the four row quarters via shared memory
#pragma unroll
      for (int ni = 0; ni < 2; ni++)
#pragma unroll
        for (int j = 0; j < 2; j++) {
          double v = LINK == 0 ? (ni < nb ? logit_acc_value(lacc[ni][j], n_pad) : 0.0) : lpacc[ni][j];
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          lpacc[ni][j] = v;
        }
      if (gid == 0) {
#pragma unroll
        for (int ni = 0; ni < 2; ni++)
#pragma unroll
          for (int j = 0; j < 2; j++) lpred[a_rq * TC + a_cb + ni * 8 + 2 * tig + j] = lpacc[ni][j];
      }
      group_bar_sync(bar_id, 256);
      if (gtid < 32 && slot0 + grp * 32 + gtid < n_list) {
        const int c = grp * 32 + gtid;
        a.LPp[(size_t)p * a.cap + slot0 + c] = ((lpred[c] + lpred[TC + c]) + lpred[2 * TC + c]) + lpred[3 * TC + c];
      }
      group_bar_sync(bar_id, 256);  // lpred (our columns) is read before the next chain tile's partials are written
      rt = 0;
      ct++;
    } else {
      rt++;
    }
  }
}

size_t glm_tc6_smem_bytes(int Dk) {
  return ((size_t)TC * xs_ld_for(Dk) + (size_t)2 * TR * xs_ld_for(Dk) + (size_t)TR * DS_LD + 2 * TR + 4 * TC + 32 + 3 + 2) * sizeof(double);
}


// =====================================================================================================================
// v7 (round 2): ONE warp per SM sub-partition, the link evaluated in the shadow of the tensor work.
// What the r02d / r02h measurements say: (i) a single warp per sub-partition issuing DMMAs back to back reaches 93 % of the FP64
// tensor pipe (tools/microbench/dmma_shapes.cu: 34.1 of 37.1 TFLOP/s) -- the pipe does not need several warps, it needs a warp
// that always has a DMMA to issue; (ii) v2's tensor phases are pipe-bound but are only 62 % of its time: the rest is link,
// block-wide barriers and tile set-up, during which all eight warps leave the pipe idle together; (iii) splitting the CTA into
// groups (v5, v6) does not help because the groups still meet at the shared tile buffers.
// So: four warps per CTA, each owning 16 chains of the chain tile for ALL rows of a (32-row) tile -- phase B's operand is then
// produced by the warp itself (a private derivative tile in shared memory, __syncwarp instead of barriers), nothing is shared
// between warps but the read-only X / q tiles -- and inside one warp the tensor work never pauses: iteration k issues phase A
// of tile k + 1 (208 DMMAs), then phase B of tile k (208 DMMAs) with the link of tile k + 1 woven into its eight r-steps, two
// elements per step, so the link's scalar instructions issue in the 15 free slots behind every DMMA.  X tiles are three TMA
// stages deep (tile k is released after phase B of tile k, and tile k + 3 has a whole iteration to arrive).
// 16 chains per warp keep v2's operand reuse (0.75 / 0.58 fragment loads per DMMA; v3 / v4 had 8 chains per warp and 1.1).
// Summation order per chain: a lane adds its rows tile by tile, the eight row-lanes are combined by shuffles -- not v2's bits,
// fixed all the same.
// =====================================================================================================================
constexpr int NT7 = 128, TR7 = 32, NST7 = 3, DW7_LD = 20;   // 4 warps; 32-row tiles; 3 X stages; leading dimension of a warp's D tile (== 4 mod 8)

__device__ __forceinline__ bool warp_arrive_is_last4(int* cnt, int lane) {
  __syncwarp();
  int last = 0;
  if (lane == 0) {
    __threadfence_block();
    last = atomicAdd(cnt, 1) == NT7 / 32 - 1;
    if (last) {
      *cnt = 0;
      __threadfence_block();
    }
  }
  return __shfl_sync(0xffffffffu, last, 0) != 0;
}

template <int LINK, int NMB>   // NMB: 8-row blocks of G (Dk <= 8 NMB)
__global__ void __launch_bounds__(NT7, 1) glm_dmma7_kernel(const GlmTcArgs a) {
  B200_DYN_SMEM(smem);
  const int Dk = a.Dk;
  const int xs_ld = xs_ld_for(Dk);
  double* Bs = smem;                                   // [TC][xs_ld]          q rows of the chain tile
  double* Xs0 = Bs + (size_t)TC * xs_ld;               // [NST7][TR7][xs_ld]   X tiles
  double* Dw0 = Xs0 + (size_t)NST7 * TR7 * xs_ld;      // [4 warps][2][TR7][DW7_LD]  each warp's derivative tiles (double-buffered)
  double* ys0 = Dw0 + (size_t)4 * 2 * TR7 * DW7_LD;    // [NST7][TR7]
  double* exptab = ys0 + NST7 * TR7;                   // [32]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(exptab + 32);  // fullX[NST7], fullB
  unsigned long long* fullX = bars;
  unsigned long long* fullB = bars + NST7;
  int* doneX = reinterpret_cast<int*>(bars + NST7 + 1);   // [NST7]
  int* doneB = doneX + NST7;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int p = blockIdx.x;
  const int n_list = a.n_list ? *a.n_list : a.n_fixed;
  const int n_ct = (n_list + TC - 1) / TC;

  // fixed row group of this CTA, in tiles of TR7 rows (a function of N only)
  const long long n_row_tiles = ((long long)a.N + TR7 - 1) / TR7;
  const long long t_begin = (long long)p * n_row_tiles / a.P, t_end = (long long)(p + 1) * n_row_tiles / a.P;
  const long long row_begin = t_begin * TR7;
  long long row_end = t_end * TR7;
  if (row_end > a.N) row_end = a.N;
  const int n_rt = (int)(t_end - t_begin);
  const long long n_it = (long long)n_ct * n_rt;
  if (n_it == 0) return;

  if (tid == 0) {
    for (int b = 0; b < NST7; b++) {
      mbar_init(&fullX[b], 1);
      doneX[b] = 0;
    }
    mbar_init(fullB, 1);
    doneB[0] = 0;
    fence_mbarrier_init();
  }
  exp_table_load(exptab, tid);
  for (int e = tid; e < TC * (Dk - a.ldq); e += NT7) {  // columns [ldq, Dk) of the q tile are never written by the bulk copies
    const int c = e / (Dk - a.ldq), k = a.ldq + e - c * (Dk - a.ldq);
    Bs[c * xs_ld + k] = 0.0;
  }
  fence_proxy_async();
  __syncthreads();  // the only block-wide barrier of the kernel

  const unsigned row_bytes = (unsigned)Dk * 8u, q_bytes = (unsigned)a.ldq * 8u;
  auto load_x = [&](long long it2) {  // one whole warp: X tile + y slice of iteration it2 into stage it2 % NST7
    const int buf = (int)(it2 % NST7);
    const int rt2 = (int)(it2 % n_rt);
    double* Xs = Xs0 + (size_t)buf * TR7 * xs_ld;
    const long long row0 = row_begin + (long long)rt2 * TR7;
    const int valid = (int)((row_end - row0) < TR7 ? (row_end - row0) : TR7);
    {
      const int r = lane;
      const int yv = r < valid ? a.y[row0 + r] : 0;
      ys0[buf * TR7 + r] = LINK == 0 ? 2.0 * yv - 1.0 : (double)yv;
      if (r >= valid)
        for (int k = 0; k < Dk; k += 2) *reinterpret_cast<double2*>(Xs + r * xs_ld + k) = make_double2(0.0, 0.0);
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_expect_tx(&fullX[buf], (unsigned)valid * row_bytes);
    __syncwarp();
    if (lane < valid) tma_load_1d(Xs + lane * xs_ld, a.X + (size_t)(row0 + lane) * a.ldx, row_bytes, &fullX[buf]);
  };
  auto load_b = [&](int ct2) {  // one whole warp: q rows of chain tile ct2
    const int nbq = n_list - ct2 * TC < TC ? n_list - ct2 * TC : TC;
    fence_proxy_async();
    if (lane == 0) mbar_expect_tx(fullB, (unsigned)nbq * q_bytes);
    __syncwarp();
#pragma unroll
    for (int h = 0; h < TC / 32; h++) {
      const int c = lane + 32 * h;
      if (c < nbq) {
        const int slot = ct2 * TC + c;
        tma_load_1d(Bs + c * xs_ld, a.Q + (size_t)(a.list ? a.list[slot] : slot) * a.ldq, q_bytes, fullB);
      }
    }
  };
  if (warp == 0) {
    load_b(0);
    for (int b = 0; b < NST7; b++)
      if (b < n_it) load_x(b);
  }

  const int cb = warp * 16;                                  // this warp's chains inside the chain tile
  double* Dw = Dw0 + (size_t)warp * 2 * TR7 * DW7_LD;
  double gacc[NMB][2][2];
  double acc[TR7 / 8][2][2];                                 // Eta of the tile whose link is pending
  double lpacc[2][2];
  LogitAcc lacc[2][2];
  int n_pad = 0;
  unsigned xphase = 0;                                       // bit b: parity of the next phase of fullX[b] this warp waits for
#pragma unroll
  for (int m = 0; m < NMB; m++)
#pragma unroll
    for (int n = 0; n < 2; n++) gacc[m][n][0] = gacc[m][n][1] = 0.0;
  lpacc[0][0] = lpacc[0][1] = lpacc[1][0] = lpacc[1][1] = 0.0;
#pragma unroll
  for (int n = 0; n < 2; n++) {
    logit_acc_init(lacc[n][0]);
    logit_acc_init(lacc[n][1]);
  }

  // phase A of iteration `it` (chain tile ctA, row tile rtA): acc <- X tile . q columns of this warp
  auto phase_a7 = [&](long long it, int ctA, int rtA, int nbA) {
    const int buf = (int)(it % NST7);
    if (rtA == 0) mbar_wait(fullB, (unsigned)(ctA & 1));
    mbar_wait(&fullX[buf], (xphase >> buf) & 1u);
    xphase ^= 1u << buf;
#pragma unroll
    for (int mi = 0; mi < TR7 / 8; mi++)
#pragma unroll
      for (int ni = 0; ni < 2; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    if (nbA > 0) {
      const double* xa = Xs0 + (size_t)buf * TR7 * xs_ld + gid * xs_ld + tig;
      const double* ba = Bs + (cb + gid) * xs_ld + tig;
#pragma unroll 2
      for (int k0 = 0; k0 < Dk; k0 += 4) {
        double av[TR7 / 8];
#pragma unroll
        for (int mi = 0; mi < TR7 / 8; mi++) av[mi] = xa[mi * 8 * xs_ld + k0];
        const double bv0 = ba[k0], bv1 = nbA == 2 ? ba[8 * xs_ld + k0] : 0.0;
#pragma unroll
        for (int mi = 0; mi < TR7 / 8; mi++) {
          dmma(acc[mi][0][0], acc[mi][0][1], av[mi], bv0);
          if (nbA == 2) dmma(acc[mi][1][0], acc[mi][1][1], av[mi], bv1);
        }
      }
    }
    if (rtA == n_rt - 1 && ctA + 1 < n_ct) {  // last row tile of the chain tile: this warp is done with the q tile
      if (warp_arrive_is_last4(doneB, lane)) load_b(ctA + 1);
    }
  };
  // link of the pending tile, elements (mi, ni): acc -> derivative, stored to the warp's D tile `dbuf`
  auto link7 = [&](int mi, int ni, int nbL, const double* ysL, long long row0L, double* DwL) {
    if (ni >= nbL) return;
    const int r = mi * 8 + gid;
    const double yd = ysL[r];
    double dv[2];
    if (LINK == 0) {
#pragma unroll
      for (int j = 0; j < 2; j++) dv[j] = logit_term(lacc[ni][j], acc[mi][ni][j], yd, exptab);
    } else {
      const bool valid = row0L + r < row_end;
      double lv[2];
#pragma unroll
      for (int j = 0; j < 2; j++) link_eval<LINK>(acc[mi][ni][j], yd, lv[j], dv[j]);
#pragma unroll
      for (int j = 0; j < 2; j++) lpacc[ni][j] += valid ? lv[j] : 0.0;
    }
    *reinterpret_cast<double2*>(&DwL[r * DW7_LD + ni * 8 + 2 * tig]) = make_double2(dv[0], dv[1]);
  };
  auto pad_rows = [&](long long row0L) {   // rows of the tile past the end of the row group this lane evaluates (LINK 0 bookkeeping)
    int n = 0;
#pragma unroll
    for (int mi = 0; mi < TR7 / 8; mi++) n += (row0L + mi * 8 + gid < row_end) ? 0 : 1;
    return n;
  };
  // partial log-likelihood of the chain tile `ctL` whose link just finished: rows are all in this warp
  auto flush_lp = [&](int ctL, int nbL) {
#pragma unroll
    for (int ni = 0; ni < 2; ni++)
#pragma unroll
      for (int j = 0; j < 2; j++) {
        double v = LINK == 0 ? (ni < nbL ? logit_acc_value(lacc[ni][j], n_pad) : 0.0) : lpacc[ni][j];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        const int slot = ctL * TC + cb + ni * 8 + 2 * tig + j;
        if (gid == 0 && slot < n_list) a.LPp[(size_t)p * a.cap + slot] = v;
        lpacc[ni][j] = 0.0;
        logit_acc_init(lacc[ni][j]);
      }
    n_pad = 0;
  };

  // ---- prologue: phase A and link of iteration 0
  int ctA = 0, rtA = 0;   // (chain tile, row tile) of the tile phase A / the link work on; phase B runs one tile behind
  {
    const int in_tile = n_list - ctA * TC;
    const int nbA = in_tile - cb > 8 ? 2 : (in_tile - cb > 0 ? 1 : 0);
    phase_a7(0, ctA, rtA, nbA);
    const long long row0L = row_begin + (long long)rtA * TR7;
    if (LINK == 0) n_pad += pad_rows(row0L);
#pragma unroll
    for (int s = 0; s < TR7 / 4; s++) link7(s >> 1, s & 1, nbA, ys0 + 0 * TR7, row0L, Dw);
    if (LINK == 0) {
#pragma unroll
      for (int ni = 0; ni < 2; ni++) {
        logit_acc_fold(lacc[ni][0]);
        logit_acc_fold(lacc[ni][1]);
      }
    }
    __syncwarp();
  }
  for (long long it = 0; it < n_it; it++) {
    // tile `it` = (ctB, rtB): its derivative tile is in Dw[it & 1]; tile it + 1 = (ctA', rtA') is next for phase A + link
    const int ctB = ctA, rtB = rtA;
    const int in_tileB = n_list - ctB * TC;
    const int nbB = in_tileB - cb > 8 ? 2 : (in_tileB - cb > 0 ? 1 : 0);
    const bool more = it + 1 < n_it;
    int nbA = 0;
    long long row0L = 0;
    if (more) {
      rtA = rtB + 1 < n_rt ? rtB + 1 : 0;
      ctA = rtA == 0 ? ctB + 1 : ctB;
      if (rtA == 0) flush_lp(ctB, nbB);   // the link of chain tile ctB is complete (it ran ahead of phase B)
      const int in_tileA = n_list - ctA * TC;
      nbA = in_tileA - cb > 8 ? 2 : (in_tileA - cb > 0 ? 1 : 0);
      phase_a7(it + 1, ctA, rtA, nbA);
      row0L = row_begin + (long long)rtA * TR7;
      if (LINK == 0) n_pad += pad_rows(row0L);
    }
    // ---- phase B of tile `it` with the link of tile it + 1 woven in: r-step s is followed by elements (s >> 1, s & 1)
    {
      const int bufB = (int)(it % NST7);
      const double* xb = Xs0 + (size_t)bufB * TR7 * xs_ld + tig * xs_ld + gid;
      const double* db = Dw + (size_t)(it & 1) * TR7 * DW7_LD + tig * DW7_LD + gid;
      double* DwL = Dw + (size_t)((it + 1) & 1) * TR7 * DW7_LD;
      const double* ysL = ys0 + (int)((it + 1) % NST7) * TR7;
      if (Dk == 8 * NMB && nbB == 2 && more && nbA == 2) {
        // steady state (full chain blocks, every design-column block in use): ONE basic block, so that the scheduler can place the
        // link's scalar instructions between the DMMAs
#pragma unroll
        for (int s = 0; s < TR7 / 4; s++) {
          const double bv0 = db[4 * s * DW7_LD], bv1 = db[4 * s * DW7_LD + 8];
#pragma unroll
          for (int m = 0; m < NMB; m++) {
            const double av = xb[4 * s * xs_ld + m * 8];
            dmma(gacc[m][0][0], gacc[m][0][1], av, bv0);
            dmma(gacc[m][1][0], gacc[m][1][1], av, bv1);
          }
          {
            const int mi = s >> 1, ni = s & 1;
            const int r = mi * 8 + gid;
            const double yd = ysL[r];
            double dv[2];
            if (LINK == 0) {
#pragma unroll
              for (int j = 0; j < 2; j++) dv[j] = logit_term(lacc[ni][j], acc[mi][ni][j], yd, exptab);
            } else {
              const bool valid = row0L + r < row_end;
              double lv[2];
#pragma unroll
              for (int j = 0; j < 2; j++) link_eval<LINK>(acc[mi][ni][j], yd, lv[j], dv[j]);
#pragma unroll
              for (int j = 0; j < 2; j++) lpacc[ni][j] += valid ? lv[j] : 0.0;
            }
            *reinterpret_cast<double2*>(&DwL[r * DW7_LD + ni * 8 + 2 * tig]) = make_double2(dv[0], dv[1]);
          }
        }
      } else
#pragma unroll
      for (int s = 0; s < TR7 / 4; s++) {
        if (nbB > 0) {
          const double bv0 = db[4 * s * DW7_LD], bv1 = nbB == 2 ? db[4 * s * DW7_LD + 8] : 0.0;
#pragma unroll
          for (int m = 0; m < NMB; m++) {
            if (m * 8 < Dk) {  // warp-uniform
              const double av = xb[4 * s * xs_ld + m * 8];
              dmma(gacc[m][0][0], gacc[m][0][1], av, bv0);
              if (nbB == 2) dmma(gacc[m][1][0], gacc[m][1][1], av, bv1);
            }
          }
        }
        if (more) link7(s >> 1, s & 1, nbA, ysL, row0L, DwL);
      }
      if (more && LINK == 0) {
#pragma unroll
        for (int ni = 0; ni < 2; ni++) {
          logit_acc_fold(lacc[ni][0]);
          logit_acc_fold(lacc[ni][1]);
        }
      }
    }
    // this warp is done with X stage `it % NST7`: the last of the four warps to get here refills it for iteration it + NST7
    if (it + NST7 < n_it) {
      if (warp_arrive_is_last4(&doneX[it % NST7], lane)) load_x(it + NST7);
    } else {
      __syncwarp();
    }
    if (rtB == n_rt - 1) {  // partial gradient of this (row group, chain tile)
      const int slot0 = ctB * TC + cb;
#pragma unroll
      for (int m = 0; m < NMB; m++) {
        const int k = m * 8 + gid;
        if (k < Dk) {
#pragma unroll
          for (int ni = 0; ni < 2; ni++)
#pragma unroll
            for (int j = 0; j < 2; j++) {
              const int slot = slot0 + ni * 8 + 2 * tig + j;
              if (slot < n_list) a.Gp[((size_t)p * a.cap + slot) * Dk + k] = gacc[m][ni][j];
              gacc[m][ni][j] = 0.0;
            }
        }
      }
      if (!more) flush_lp(ctB, nbB);
    }
  }
}

size_t glm_tc7_smem_bytes(int Dk) {
  return ((size_t)TC * xs_ld_for(Dk) + (size_t)NST7 * TR7 * xs_ld_for(Dk) + (size_t)4 * 2 * TR7 * DW7_LD + NST7 * TR7 + 32 + NST7 + 1 + NST7) * sizeof(double);
}


// =====================================================================================================================
// v3: the same fused product, with NO block-wide barrier in steady state (profiles/README.md r01c: with three
// __syncthreads per tile all eight warps run the same phase at the same time, the link phase cannot saturate the shared
// FP64 pipe on its own, and 24 % of its cycles idle).
//
//  * warp w (8 of them) owns the 8 chains [8w, 8w+8) of the chain tile for ALL 64 rows of the row tile: phase A 8 row
//    blocks x 1 chain block, link on its own accumulators, derivative tile Dw[64][8] in its private slice of shared
//    memory, phase B Dk/8 design-column blocks x 1 chain block.  Its phase-B operand is produced by itself, so a
//    __syncwarp replaces the __syncthreads; the log-likelihood partial needs only shuffles;
//  * nobody waits to refill a buffer: a warp that is done with X buffer `buf` bumps a shared counter, and the LAST warp to
//    do so issues the TMA bulk copies of the tile two iterations ahead into that buffer (same for the q tile of the next
//    chain tile after phase A of the last row tile).  Consumers only ever wait on the "full" mbarriers;
//  * warps therefore drift apart by up to one tile: while one warp of an SM sub-partition evaluates the link function
//    (latency-bound scalar FP64), the other one issues DMMAs.
// Row groups, partial buffers and summation order per chain are unchanged (fixed order; independent of the batch).
// =====================================================================================================================
constexpr int NCW = 8;                 // warps per CTA
constexpr int NT3 = 32 * NCW;
constexpr int DW_LD = 12;              // leading dimension of a warp's derivative tile [TR][8] (== 4 mod 8: conflict-free fragments)

// one arrival per warp on a shared counter; returns true (to the whole warp) in the warp that arrived last, after
// resetting the counter.  Orders the warp's earlier shared-memory reads before the refill the last arriver issues.
__device__ __forceinline__ bool warp_arrive_is_last(int* cnt, int lane) {
  __syncwarp();
  int last = 0;
  if (lane == 0) {
    __threadfence_block();
    last = atomicAdd(cnt, 1) == NCW - 1;
    if (last) {
      *cnt = 0;
      __threadfence_block();
    }
  }
  return __shfl_sync(0xffffffffu, last, 0) != 0;
}

template <int LINK, int MB>
__global__ void __launch_bounds__(NT3, 1) glm_dmma3_kernel(const GlmTcArgs a) {
  B200_DYN_SMEM(smem);
  const int Dk = a.Dk;
  const int xs_ld = xs_ld_for(Dk);
  double* Bs = smem;                                // [TC][xs_ld]
  double* Xs0 = Bs + (size_t)TC * xs_ld;            // [2][TR][xs_ld]
  double* Dw0 = Xs0 + (size_t)2 * TR * xs_ld;       // [NCW][TR][DW_LD]
  double* ys0 = Dw0 + (size_t)NCW * TR * DW_LD;     // [2][TR]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(ys0 + 2 * TR);
  unsigned long long* fullX = bars;        // [2]
  unsigned long long* fullB = bars + 2;
  int* doneX = reinterpret_cast<int*>(bars + 3);  // [2] warps done with X buffer b;  [2]: warps done with the q tile
  int* doneB = doneX + 2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int p = blockIdx.x;
  const int n_list = a.n_list ? *a.n_list : a.n_fixed;
  const int n_ct = (n_list + TC - 1) / TC;

  const long long t_begin = (long long)p * a.n_row_tiles / a.P, t_end = (long long)(p + 1) * a.n_row_tiles / a.P;
  const long long row_begin = t_begin * TR;
  long long row_end = t_end * TR;
  if (row_end > a.N) row_end = a.N;
  const int n_rt = (int)(t_end - t_begin);
  const long long n_it = (long long)n_ct * n_rt;
  if (n_it == 0) return;

  if (tid == 0) {
    mbar_init(&fullX[0], 1);
    mbar_init(&fullX[1], 1);
    mbar_init(fullB, 1);
    doneX[0] = doneX[1] = doneB[0] = 0;
    fence_mbarrier_init();
  }
  for (int e = tid; e < TC * (Dk - a.ldq); e += NT3) {  // columns [ldq, Dk) of the q tile are never written by the bulk copies
    const int c = e / (Dk - a.ldq), k = a.ldq + e - c * (Dk - a.ldq);
    Bs[c * xs_ld + k] = 0.0;
  }
  fence_proxy_async();
  __syncthreads();  // the only block-wide barrier of the kernel

  const unsigned row_bytes = (unsigned)Dk * 8u, q_bytes = (unsigned)a.ldq * 8u;
  // issued by one whole warp: X tile + y slice of iteration `it2` into its buffer
  auto load_x = [&](long long it2) {
    const int buf = (int)(it2 & 1);
    const int rt2 = (int)(it2 % n_rt);
    double* Xs = Xs0 + (size_t)buf * TR * xs_ld;
    const long long row0 = row_begin + (long long)rt2 * TR;
    const int valid = (int)((row_end - row0) < TR ? (row_end - row0) : TR);
    // generic-proxy part of the tile first (y slice, zero rows past the end of the group): released by the arrive below
#pragma unroll
    for (int h = 0; h < TR / 32; h++) {
      const int r = lane + 32 * h;
      const int yv = r < valid ? a.y[row0 + r] : 0;
      ys0[buf * TR + r] = LINK == 0 ? 2.0 * yv - 1.0 : (double)yv;
      if (r >= valid)
        for (int k = 0; k < Dk; k += 2) *reinterpret_cast<double2*>(Xs + r * xs_ld + k) = make_double2(0.0, 0.0);
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_expect_tx(&fullX[buf], (unsigned)valid * row_bytes);
    __syncwarp();
#pragma unroll
    for (int h = 0; h < TR / 32; h++) {
      const int r = lane + 32 * h;
      if (r < valid) tma_load_1d(Xs + r * xs_ld, a.X + (size_t)(row0 + r) * a.ldx, row_bytes, &fullX[buf]);
    }
  };
  auto load_b = [&](int ct2) {  // one whole warp: q rows of chain tile ct2
    const int nb = n_list - ct2 * TC < TC ? n_list - ct2 * TC : TC;
    fence_proxy_async();
    if (lane == 0) mbar_expect_tx(fullB, (unsigned)nb * q_bytes);
    __syncwarp();
#pragma unroll
    for (int h = 0; h < TC / 32; h++) {
      const int c = lane + 32 * h;
      if (c < nb) {
        const int slot = ct2 * TC + c;
        tma_load_1d(Bs + c * xs_ld, a.Q + (size_t)(a.list ? a.list[slot] : slot) * a.ldq, q_bytes, fullB);
      }
    }
  };
  if (warp == 0) {
    load_b(0);
    load_x(0);
    if (n_it > 1) load_x(1);
  }

  const int cb = warp * 8;  // this warp's chains inside the chain tile
  double* Dw = Dw0 + (size_t)warp * TR * DW_LD;
  double gacc[MB][2];
  double lpacc[2];
  int ct = 0, rt = 0;
  for (long long it = 0; it < n_it; it++) {
    const int buf = (int)(it & 1);
    const bool have = n_list - ct * TC > cb;  // warp-uniform: does this warp's block hold requesting chains?
    if (rt == 0) {
#pragma unroll
      for (int m = 0; m < MB; m++) gacc[m][0] = gacc[m][1] = 0.0;
      lpacc[0] = lpacc[1] = 0.0;
      mbar_wait(fullB, (unsigned)(ct & 1));
    }
    mbar_wait(&fullX[buf], (unsigned)((it >> 1) & 1));
    const double* Xs = Xs0 + (size_t)buf * TR * xs_ld;
    const double* ys = ys0 + buf * TR;
    const long long row0 = row_begin + (long long)rt * TR;
    double acc[8][2];
    if (have) {
      // ---- phase A: Eta[64 rows][8 chains]
#pragma unroll
      for (int mi = 0; mi < 8; mi++) acc[mi][0] = acc[mi][1] = 0.0;
      const double* xa = Xs + gid * xs_ld + tig;
      const double* ba = Bs + (cb + gid) * xs_ld + tig;
#pragma unroll 2
      for (int k0 = 0; k0 < Dk; k0 += 4) {
        double av[8];
#pragma unroll
        for (int mi = 0; mi < 8; mi++) av[mi] = xa[mi * 8 * xs_ld + k0];
        const double bv = ba[k0];
#pragma unroll
        for (int mi = 0; mi < 8; mi++) dmma(acc[mi][0], acc[mi][1], av[mi], bv);
      }
    }
    if (rt == n_rt - 1 && ct + 1 < n_ct) {  // last row tile of this chain tile: this warp is done with the q tile
      if (warp_arrive_is_last(doneB, lane)) load_b(ct + 1);
    }
    if (have) {
      // ---- link on the accumulator fragments: element (row = mi*8 + gid, chain = cb + 2*tig + j)
#pragma unroll
      for (int mi = 0; mi < 8; mi++) {
        const int r = mi * 8 + gid;
        const bool valid = row0 + r < row_end;
        const double yd = ys[r];
        double lv[2], dv[2];
#pragma unroll
        for (int j = 0; j < 2; j++) link_eval<LINK>(acc[mi][j], yd, lv[j], dv[j]);
#pragma unroll
        for (int j = 0; j < 2; j++) lpacc[j] += valid ? lv[j] : 0.0;
        *reinterpret_cast<double2*>(&Dw[r * DW_LD + 2 * tig]) = make_double2(dv[0], dv[1]);
      }
      __syncwarp();
      // ---- phase B: G[Dk][8 chains] += X tile^T . Dw
      const double* xb = Xs + tig * xs_ld + gid;
      const double* db = Dw + tig * DW_LD + gid;
#pragma unroll 2
      for (int r0 = 0; r0 < TR; r0 += 4) {
        const double bv = db[r0 * DW_LD];
#pragma unroll
        for (int m = 0; m < MB; m++) {
          if (m * 8 < Dk) {  // warp-uniform
            const double av = xb[r0 * xs_ld + m * 8];
            dmma(gacc[m][0], gacc[m][1], av, bv);
          }
        }
      }
    }
    // this warp is done with Xs[buf] (and with its Dw): the last warp to get here refills the buffer for iteration it + 2
    if (it + 2 < n_it) {
      if (warp_arrive_is_last(&doneX[buf], lane)) load_x(it + 2);
    } else {
      __syncwarp();
    }
    if (rt == n_rt - 1) {
      if (have) {
        const int slot0 = ct * TC + cb;
#pragma unroll
        for (int m = 0; m < MB; m++) {
          const int k = m * 8 + gid;
          if (k < Dk) {
#pragma unroll
            for (int j = 0; j < 2; j++) {
              const int slot = slot0 + 2 * tig + j;
              if (slot < n_list) a.Gp[((size_t)p * a.cap + slot) * Dk + k] = gacc[m][j];
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 2; j++) {  // sum over the 8 row lanes (gid) in a fixed order
          double v = lpacc[j];
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          const int slot = slot0 + 2 * tig + j;
          if (gid == 0 && slot < n_list) a.LPp[(size_t)p * a.cap + slot] = v;
        }
      }
      rt = 0;
      ct++;
    } else {
      rt++;
    }
  }
}


// =====================================================================================================================
// v4 (EXPERIMENTAL, not validated on hardware yet: selected only by B200NUTS_GLM_VARIANT=3, covered by the GPU tests only
// when B200_TEST_EXPERIMENTAL=1): v3 with the two changes its r01d profile asked for.
//  * the derivative tile never touches shared memory: after the link a lane holds D[row mi*8+gid][chain 2*tig+j] in its
//    accumulator registers; the phase-B operand fragment b0 = D[row r0+tig][chain gid] is fetched from the owning lane with
//    two 64-bit shuffles (one per j, the destination selects) -- 64 shuffles per warp and tile in place of 49 KB of smem;
//  * the freed shared memory holds a THIRD X stage (Dk <= 104): the refill of a buffer is still issued by the last warp to
//    release it, but it is needed two tiles later instead of one, so warp drift no longer turns into mbarrier spinning.
// Summation order per chain is that of v3.
// =====================================================================================================================
template <int LINK, int MB, int NST>
__global__ void __launch_bounds__(NT3, 1) glm_dmma4_kernel(const GlmTcArgs a) {
  B200_DYN_SMEM(smem);
  const int Dk = a.Dk;
  const int xs_ld = xs_ld_for(Dk);
  double* Bs = smem;                                  // [TC][xs_ld]
  double* Xs0 = Bs + (size_t)TC * xs_ld;              // [NST][TR][xs_ld]
  double* ys0 = Xs0 + (size_t)NST * TR * xs_ld;       // [NST][TR]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(ys0 + NST * TR);
  unsigned long long* fullX = bars;                   // [NST]
  unsigned long long* fullB = bars + NST;
  int* doneX = reinterpret_cast<int*>(bars + NST + 1);  // [NST] warps done with X buffer b
  int* doneB = doneX + NST;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int p = blockIdx.x;
  const int n_list = a.n_list ? *a.n_list : a.n_fixed;
  const int n_ct = (n_list + TC - 1) / TC;

  const long long t_begin = (long long)p * a.n_row_tiles / a.P, t_end = (long long)(p + 1) * a.n_row_tiles / a.P;
  const long long row_begin = t_begin * TR;
  long long row_end = t_end * TR;
  if (row_end > a.N) row_end = a.N;
  const int n_rt = (int)(t_end - t_begin);
  const long long n_it = (long long)n_ct * n_rt;
  if (n_it == 0) return;

  if (tid == 0) {
    for (int b = 0; b < NST; b++) {
      mbar_init(&fullX[b], 1);
      doneX[b] = 0;
    }
    mbar_init(fullB, 1);
    doneB[0] = 0;
    fence_mbarrier_init();
  }
  for (int e = tid; e < TC * (Dk - a.ldq); e += NT3) {
    const int c = e / (Dk - a.ldq), k = a.ldq + e - c * (Dk - a.ldq);
    Bs[c * xs_ld + k] = 0.0;
  }
  fence_proxy_async();
  __syncthreads();  // the only block-wide barrier of the kernel

  const unsigned row_bytes = (unsigned)Dk * 8u, q_bytes = (unsigned)a.ldq * 8u;
  auto load_x = [&](long long it2) {  // one whole warp: X tile + y slice of iteration it2 into buffer it2 % NST
    const int buf = (int)(it2 % NST);
    const int rt2 = (int)(it2 % n_rt);
    double* Xs = Xs0 + (size_t)buf * TR * xs_ld;
    const long long row0 = row_begin + (long long)rt2 * TR;
    const int valid = (int)((row_end - row0) < TR ? (row_end - row0) : TR);
#pragma unroll
    for (int h = 0; h < TR / 32; h++) {
      const int r = lane + 32 * h;
      const int yv = r < valid ? a.y[row0 + r] : 0;
      ys0[buf * TR + r] = LINK == 0 ? 2.0 * yv - 1.0 : (double)yv;
      if (r >= valid)
        for (int k = 0; k < Dk; k += 2) *reinterpret_cast<double2*>(Xs + r * xs_ld + k) = make_double2(0.0, 0.0);
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_expect_tx(&fullX[buf], (unsigned)valid * row_bytes);
    __syncwarp();
#pragma unroll
    for (int h = 0; h < TR / 32; h++) {
      const int r = lane + 32 * h;
      if (r < valid) tma_load_1d(Xs + r * xs_ld, a.X + (size_t)(row0 + r) * a.ldx, row_bytes, &fullX[buf]);
    }
  };
  auto load_b = [&](int ct2) {  // one whole warp: q rows of chain tile ct2
    const int nb = n_list - ct2 * TC < TC ? n_list - ct2 * TC : TC;
    fence_proxy_async();
    if (lane == 0) mbar_expect_tx(fullB, (unsigned)nb * q_bytes);
    __syncwarp();
#pragma unroll
    for (int h = 0; h < TC / 32; h++) {
      const int c = lane + 32 * h;
      if (c < nb) {
        const int slot = ct2 * TC + c;
        tma_load_1d(Bs + c * xs_ld, a.Q + (size_t)(a.list ? a.list[slot] : slot) * a.ldq, q_bytes, fullB);
      }
    }
  };
  if (warp == 0) {
    load_b(0);
    for (int b = 0; b < NST; b++)
      if (b < n_it) load_x(b);
  }

  const int cb = warp * 8;
  double gacc[MB][2];
  double lpacc[2];
  int ct = 0, rt = 0;
  int buf = 0;           // it % NST
  unsigned xphase = 0;   // bit b: parity of the next phase of fullX[b] this warp waits for
  for (long long it = 0; it < n_it; it++) {
    const bool have = n_list - ct * TC > cb;
    if (rt == 0) {
#pragma unroll
      for (int m = 0; m < MB; m++) gacc[m][0] = gacc[m][1] = 0.0;
      lpacc[0] = lpacc[1] = 0.0;
      mbar_wait(fullB, (unsigned)(ct & 1));
    }
    mbar_wait(&fullX[buf], (xphase >> buf) & 1u);
    xphase ^= 1u << buf;
    const double* Xs = Xs0 + (size_t)buf * TR * xs_ld;
    const double* ys = ys0 + buf * TR;
    const long long row0 = row_begin + (long long)rt * TR;
    double acc[8][2];  // Eta, then the derivative tile
    if (have) {
#pragma unroll
      for (int mi = 0; mi < 8; mi++) acc[mi][0] = acc[mi][1] = 0.0;
      const double* xa = Xs + gid * xs_ld + tig;
      const double* ba = Bs + (cb + gid) * xs_ld + tig;
#pragma unroll 2
      for (int k0 = 0; k0 < Dk; k0 += 4) {
        double av[8];
#pragma unroll
        for (int mi = 0; mi < 8; mi++) av[mi] = xa[mi * 8 * xs_ld + k0];
        const double bv = ba[k0];
#pragma unroll
        for (int mi = 0; mi < 8; mi++) dmma(acc[mi][0], acc[mi][1], av[mi], bv);
      }
    }
    if (rt == n_rt - 1 && ct + 1 < n_ct) {
      if (warp_arrive_is_last(doneB, lane)) load_b(ct + 1);
    }
    if (have) {
      // ---- link in place: acc[mi][j] becomes D[row mi*8 + gid][chain cb + 2*tig + j]
#pragma unroll
      for (int mi = 0; mi < 8; mi++) {
        const int r = mi * 8 + gid;
        const bool valid = row0 + r < row_end;
        const double yd = ys[r];
        double lv[2], dv[2];
#pragma unroll
        for (int j = 0; j < 2; j++) link_eval<LINK>(acc[mi][j], yd, lv[j], dv[j]);
#pragma unroll
        for (int j = 0; j < 2; j++) {
          lpacc[j] += valid ? lv[j] : 0.0;
          acc[mi][j] = dv[j];
        }
      }
      // ---- phase B: G[Dk][8 chains] += X tile^T . D, the D fragment fetched from its owner lane by shuffles
      const double* xb = Xs + tig * xs_ld + gid;
      const int sel = gid & 1;
#pragma unroll
      for (int s = 0; s < TR / 4; s++) {   // rows r0 = 4 s .. 4 s + 3 = 8 (s >> 1) + 4 (s & 1) + tig
        const int src = (((s & 1) * 4 + tig) << 2) | (gid >> 1);
        const double v0 = __shfl_sync(0xffffffffu, acc[s >> 1][0], src);
        const double v1 = __shfl_sync(0xffffffffu, acc[s >> 1][1], src);
        const double bv = sel ? v1 : v0;
#pragma unroll
        for (int m = 0; m < MB; m++) {
          if (m * 8 < Dk) {  // warp-uniform
            const double av = xb[(4 * s) * xs_ld + m * 8];
            dmma(gacc[m][0], gacc[m][1], av, bv);
          }
        }
      }
    }
    // this warp is done with X buffer `buf`: the last warp to get here refills it for iteration it + NST
    if (it + NST < n_it) {
      if (warp_arrive_is_last(&doneX[buf], lane)) load_x(it + NST);
    } else {
      __syncwarp();
    }
    if (rt == n_rt - 1) {
      if (have) {
        const int slot0 = ct * TC + cb;
#pragma unroll
        for (int m = 0; m < MB; m++) {
          const int k = m * 8 + gid;
          if (k < Dk) {
#pragma unroll
            for (int j = 0; j < 2; j++) {
              const int slot = slot0 + 2 * tig + j;
              if (slot < n_list) a.Gp[((size_t)p * a.cap + slot) * Dk + k] = gacc[m][j];
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 2; j++) {
          double v = lpacc[j];
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          const int slot = slot0 + 2 * tig + j;
          if (gid == 0 && slot < n_list) a.LPp[(size_t)p * a.cap + slot] = v;
        }
      }
      rt = 0;
      ct++;
    } else {
      rt++;
    }
    buf = buf + 1 == NST ? 0 : buf + 1;
  }
}

size_t glm_tc4_smem_bytes(int Dk, int nst) {
  return ((size_t)TC * xs_ld_for(Dk) + (size_t)nst * TR * xs_ld_for(Dk) + (size_t)nst * TR + 2 * nst + 4) * sizeof(double);
}

size_t glm_tc3_smem_bytes(int Dk) {
  return ((size_t)TC * xs_ld_for(Dk) + (size_t)2 * TR * xs_ld_for(Dk) + (size_t)NCW * TR * DW_LD + 2 * TR + 8) * sizeof(double);
}

}  // namespace

int glm_tc_max_dk() { return 112; }

size_t glm_tc_smem_bytes(int Dk) {
  return ((size_t)TC * xs_ld_for(Dk) + (size_t)2 * TR * xs_ld_for(Dk) + (size_t)TR * DS_LD + 2 * TR + 4 + 32) * sizeof(double);
}

int launch_glm_dmma(const ModelData& m, const GradBatch& b, const GlmWorkspace& ws, cudaStream_t stream) {
  GlmTcArgs a;
  a.X = m.X;
  a.y = m.yi;
  a.Q = b.Q;
  a.list = b.list;
  a.n_list = b.n_list;
  a.Gp = ws.Gp;
  a.LPp = ws.LPp;
  a.n_row_tiles = ((long long)m.N + TR - 1) / TR;
  a.N = m.N;
  a.ldx = m.ldx;
  a.Dk = ws.Dp;
  a.D = m.D;
  a.ldq = b.ld;
  a.n_fixed = b.C;
  a.cap = ws.C;
  a.P = ws.P;
  const bool logistic = m.kind == MODEL_LOGISTIC;
  if (ws.variant == 3) {  // experimental v4: three X stages when they fit in 227 KB, else two
    const int nst = glm_tc4_smem_bytes(a.Dk, 3) <= 232448 ? 3 : 2;
    const size_t smem4 = glm_tc4_smem_bytes(a.Dk, nst);
#define LAUNCH4(LINK, MB)                                                                          \
  {                                                                                                \
    if (nst == 3) {                                                                                \
      auto k = glm_dmma4_kernel<LINK, MB, 3>;                                                      \
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4);            \
      B200_LAUNCH(k, ws.P, NT3, smem4, stream, a);                                                          \
    } else {                                                                                       \
      auto k = glm_dmma4_kernel<LINK, MB, 2>;                                                      \
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4);            \
      B200_LAUNCH(k, ws.P, NT3, smem4, stream, a);                                                          \
    }                                                                                              \
  }
    if (a.Dk <= 16) {
      if (logistic) LAUNCH4(0, 2) else LAUNCH4(1, 2)
    } else if (a.Dk <= 32) {
      if (logistic) LAUNCH4(0, 4) else LAUNCH4(1, 4)
    } else if (a.Dk <= 64) {
      if (logistic) LAUNCH4(0, 8) else LAUNCH4(1, 8)
    } else {
      if (logistic) LAUNCH4(0, 14) else LAUNCH4(1, 14)
    }
#undef LAUNCH4
    return 1;
  }
  if (ws.variant == 2) {
    const size_t smem3 = glm_tc3_smem_bytes(a.Dk);
#define LAUNCH3(LINK, MB)                                                                     \
  {                                                                                           \
    auto k = glm_dmma3_kernel<LINK, MB>;                                                      \
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3);         \
    B200_LAUNCH(k, ws.P, NT3, smem3, stream, a);                                                       \
  }
    if (a.Dk <= 16) {
      if (logistic) LAUNCH3(0, 2) else LAUNCH3(1, 2)
    } else if (a.Dk <= 32) {
      if (logistic) LAUNCH3(0, 4) else LAUNCH3(1, 4)
    } else if (a.Dk <= 64) {
      if (logistic) LAUNCH3(0, 8) else LAUNCH3(1, 8)
    } else {
      if (logistic) LAUNCH3(0, 14) else LAUNCH3(1, 14)
    }
#undef LAUNCH3
    return 1;
  }
  if (ws.variant == 6) {  // v7: one warp per sub-partition, link in the shadow of the tensor work
    const size_t smem7 = glm_tc7_smem_bytes(a.Dk);
#define LAUNCH7(LINK, NMB)                                                                    \
  {                                                                                           \
    auto k = glm_dmma7_kernel<LINK, NMB>;                                                     \
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem7);         \
    B200_LAUNCH(k, ws.P, NT7, smem7, stream, a);                                              \
  }
    if (a.Dk <= 16) {
      if (logistic) LAUNCH7(0, 2) else LAUNCH7(1, 2)
    } else if (a.Dk <= 32) {
      if (logistic) LAUNCH7(0, 4) else LAUNCH7(1, 4)
    } else if (a.Dk <= 56) {
      if (logistic) LAUNCH7(0, 7) else LAUNCH7(1, 7)
    } else if (a.Dk <= 104) {
      if (logistic) LAUNCH7(0, 13) else LAUNCH7(1, 13)
    } else {
      if (logistic) LAUNCH7(0, 14) else LAUNCH7(1, 14)
    }
#undef LAUNCH7
    return 1;
  }
  if (ws.variant == 5) {  // v6: sixteen warps, two out-of-phase groups of eight
    const size_t smem6 = glm_tc6_smem_bytes(a.Dk);
#define LAUNCH6(LINK, MB)                                                                     \
  {                                                                                           \
    auto k = glm_dmma6_kernel<LINK, MB>;                                                      \
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem6);         \
    B200_LAUNCH(k, ws.P, NT6, smem6, stream, a);                                              \
  }
    if (a.Dk <= 16) {
      if (logistic) LAUNCH6(0, 1) else LAUNCH6(1, 1)
    } else if (a.Dk <= 32) {
      if (logistic) LAUNCH6(0, 2) else LAUNCH6(1, 2)
    } else if (a.Dk <= 64) {
      if (logistic) LAUNCH6(0, 4) else LAUNCH6(1, 4)
    } else {
      if (logistic) LAUNCH6(0, 7) else LAUNCH6(1, 7)
    }
#undef LAUNCH6
    return 1;
  }
  if (ws.variant == 4) {  // v5: two out-of-phase groups of four warps
    const size_t smem5 = glm_tc5_smem_bytes(a.Dk);
#define LAUNCH5(LINK, MB)                                                                     \
  {                                                                                           \
    auto k = glm_dmma5_kernel<LINK, MB>;                                                      \
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem5);         \
    B200_LAUNCH(k, ws.P, NT, smem5, stream, a);                                               \
  }
    if (a.Dk <= 16) {
      if (logistic) LAUNCH5(0, 1) else LAUNCH5(1, 1)
    } else if (a.Dk <= 32) {
      if (logistic) LAUNCH5(0, 2) else LAUNCH5(1, 2)
    } else if (a.Dk <= 64) {
      if (logistic) LAUNCH5(0, 4) else LAUNCH5(1, 4)
    } else {
      if (logistic) LAUNCH5(0, 7) else LAUNCH5(1, 7)
    }
#undef LAUNCH5
    return 1;
  }
  const size_t smem = glm_tc_smem_bytes(a.Dk);
#define LAUNCH(LINK, MB)                                                                     \
  {                                                                                          \
    auto k = glm_dmma_kernel<LINK, MB>;                                                      \
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);         \
    B200_LAUNCH(k, ws.P, NT, smem, stream, a);                                                         \
  }
  if (a.Dk <= 16) {
    if (logistic) LAUNCH(0, 1) else LAUNCH(1, 1)
  } else if (a.Dk <= 32) {
    if (logistic) LAUNCH(0, 2) else LAUNCH(1, 2)
  } else if (a.Dk <= 64) {
    if (logistic) LAUNCH(0, 4) else LAUNCH(1, 4)
  } else {  // Dk <= 112 (larger designs use the DFMA kernel of glm.cu: the double-buffered tile would not fit in smem)
    if (logistic) LAUNCH(0, 7) else LAUNCH(1, 7)
  }
#undef LAUNCH
  return 1;
}

}  // namespace b200
