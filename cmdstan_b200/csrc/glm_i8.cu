// Ozaki-split INT8 formulation of the fused cross-chain logistic-regression gradient on the 5th-generation tensor cores
// (tcgen05.mma kind::i8, INT32 accumulators in tensor memory).  Same mathematics and the same reference semantics as glm.cu /
// glm_dmma.cu (bernoulli_logit_glm_lpmf.hpp:51-167): per row tile
//     Eta[128 x 32] = X . q ,   link in FP64 registers ,   G[Dk x 32] += X^T . dm
// but both contractions leave the FP64 pipe (tcgen05 has no FP64 kind; DMMA tops out at 37 TFLOP/s, v2 of glm_dmma.cu reaches 0.6 of
// it): every operand is a fixed-point integer V = rint(v 2^F), |v| <= 1, F = 8 NS - 2, written in NS balanced base-256 digits
// (INT8 planes; the top digit lies in [-64, 64]); the tensor cores form the EXACT integer products of digit planes and add the
// pairs (i, j) with i + j = k into one INT32 accumulator per diagonal k < NS (diagonals k >= NS, weight 2^-(8 NS) and below, are
// dropped).  NS = 6: 46-bit operands, 21 digit-plane products per contraction, lp to ~2e-14 and the gradient to ~1e-13 of its
// largest component (tools/ozaki_i8_model.py reproduces the arithmetic exactly on the CPU); NS = 5: 38 bits, 15 products, 5e-12 /
// 2e-11.  Scales: a power of two per design column (of the data, fixed), a power of two per chain for q, none for dm (|dm| <= 1).
//
// Tensor-core formulation (validated on hardware by tools/microbench/umma_i8.cu, profiles/r02m_umma_i8.txt):
//  * X is sliced ONCE per data set into the image [row tile][digit plane][row / 8][col / 16][8][16] (core matrices of 8 rows x 16
//    bytes, no swizzle).  The same shared-memory tile is the A operand of both products: K-major (M = rows, K = columns) for X . q,
//    MN-major (M = columns, K = rows) for X^T . dm -- only the descriptors differ.
//  * the digit planes of q (K-major) resp. dm (MN-major) of one chain tile are laid side by side along N, so that ONE instruction
//    per K step multiplies X plane i with planes 0 .. NS-1-i and accumulates into the column window [32 i, 32 NS) of the
//    accumulator: diagonal k = i + j lands in columns [32 k, 32 k + 32).  NS (NS + 1) / 2 products cost NS instructions per K
//    step with N = 32 (NS - i) <= 192: 84 % of the INT8 peak measured with nothing else running.
//  * roles: warp 0 streams X planes (14 KB bulk copies, a ring of >= 2 NS slots, one full / empty mbarrier pair per slot) and the
//    q planes of the chain tile; one thread of warp 1 issues every MMA and commits completions to the mbarriers; eight epilogue
//    warps (thread = row of the tile = TMEM lane, 16 chains per thread) read the Eta diagonals with tcgen05.ld, release the
//    accumulator at once (the next tile's first product runs under this tile's link), recombine to FP64, evaluate the link
//    (fastmath.cuh, 19 FP64 instructions per term), slice dm into digits on the integer pipe and store them as the MN-major B
//    operand of the second product.  G stays in tensor memory for all row tiles of a (row group, chain tile).
//  * exactness pays twice: the integer sums are associative, so the work split cannot change a bit of the gradient; only the
//    order of the FP64 log-likelihood sums has to stay a function of N alone (row groups fixed by N, as in glm_dmma.cu).
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "fastmath.cuh"
#include "glm_i8_math.cuh"
#include "kernels.h"

namespace b200 {

namespace {

constexpr int TRW = 128;      // rows per tile = M of the first product = K of the second
constexpr int CN = 32;        // chains per tile
constexpr int NTH = 576;      // warp 0: loads, warp 1: MMA issue + TMEM allocation, warps 2..17: epilogue (four per SM sub-partition)
constexpr int ECH = 8;        // chains per epilogue thread: warp group wg = (warp - 2) / 4 owns chains [8 wg, 8 wg + 8) of the tile
constexpr int NEW = 16;       // epilogue warps
constexpr int KQ = 4;         // chain-tile classes: CTA (p, kq) serves row group p and the chain tiles ct == kq (mod KQ)
// When may the next tile's first product start?  1 (shipped): as soon as the Eta accumulator has been read -- the product then
// runs under this tile's link and stretches it (scalar FP64 is starved while the tensor core runs), but its slow single-thread
// issue is off the critical path.  0: only after the link -- the FP64 phase runs undisturbed (2 x 1.45 k cycles instead of
// 3.6 k + 1.8 k) but the tile then waits for the 24 MMAs to be issued (2.2 k cycles).  Measured r03h / r03g: 0.597 vs 0.576 ms per
// c3 launch.  Both keep the integer-only recombination in front of the FP64 phase, under the previous tile's second product.
#ifndef I8_EARLY_RELEASE
#define I8_EARLY_RELEASE 1
#endif
constexpr uint32_t ACOL = 0, GCOL = 256;   // TMEM columns of the Eta / G diagonals (NS * 32 <= 192 each)
constexpr unsigned kMaxPolls = 1u << 24;   // a protocol error must end the kernel (trap), never hang the device: each try_wait suspends
                                            // the thread for a bounded time of its own, so this is seconds, not microseconds

struct GlmI8Args {
  const uint8_t* Ximg;      // [T][NS][16][DC][8][16]
  const uint8_t* Qimg;      // [chain tile][NS][4][DCQ][8][16] digits, then 32 doubles (scale of S per chain)
  const double* colscale;   // [Dk] power of two >= max |X[:, k]|
  const int* y;
  const int* n_list;
  double* Gp;               // [P][cap][Dk]
  double* LPp;              // [P][cap]
  int n_fixed, N, Dk, DC, KA, cap, P, T, NB;
  long long* trace;         // TRACE builds only: [4 roles][kTraceIt][8] clock64 stamps of CTA 0 (tools/i8_trace.py)
};
constexpr int kTraceIt = 96;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
               : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done != 0;
}
__device__ __noinline__ void mbar_timeout(uint32_t bar, uint32_t parity) {
  printf("glm_i8_kernel: mbarrier wait timed out (block %d thread %d barrier smem+%u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  unsigned polls = 0;
  while (!mbar_try(bar, parity))
    if (++polls > kMaxPolls) mbar_timeout(bar, parity);
}
// the same for the producer and the MMA thread: they wait for most of a tile's time next to four working epilogue warps of their SM
// sub-partition; a tight polling loop takes issue slots from those (and every tile waits for the slowest of the sixteen)
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  unsigned polls = 0;
  while (!mbar_try(bar, parity)) {
    __nanosleep(40);
    if (++polls > kMaxPolls) mbar_timeout(bar, parity);
  }
}
// the MMA warp waits as ONE: the vote makes the loop's exit warp-uniform, so that everything computed after it (ring cursors,
// descriptors) can live in uniform registers and an MMA costs a handful of instructions instead of an R2UR per operand word
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity) {
  unsigned polls = 0;
  while (!__all_sync(0xffffffffu, mbar_try(bar, parity))) {
    __nanosleep(40);
    if (++polls > kMaxPolls) mbar_timeout(bar, parity);
  }
}
__device__ __forceinline__ void tma_load_1d(uint32_t smem_dst, const void* gsrc, unsigned bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(bar) : "memory");
}
// the same with an L2 eviction-priority hint: the digit planes of X are re-read by every chain tile of every launch while the
// advance kernel streams the chains' state through L2 in between (r02q: 25 % of the plane reads missed L2)
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_1d_hint(uint32_t smem_dst, const void* gsrc, unsigned bytes, uint32_t bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// shared-memory matrix descriptor, no swizzle: start address, leading / stride byte offsets (all >> 4), sm_100 descriptor version
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46);
}
// instruction descriptor of kind::i8: S32 accumulators, signed x signed, operand majors (0 = K, 1 = MN), N >> 3, M >> 4
__host__ __device__ constexpr uint32_t idesc_i8(int M, int N, int mn_major) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)mn_major << 15) | ((uint32_t)mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// Issued by the whole (converged) warp 1 with one elected lane executing it: with a single active thread the compiler cannot keep
// the descriptors in uniform registers and wraps every MMA in an ELECT / R2UR.BROADCAST loop of ~25 instructions (150 cycles per
// MMA on a sub-partition shared with four working warps: r03d, 3.6 k cycles to issue the 24 MMAs of one product)
__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t@q tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// four K steps of one plane product in one block: one election, the descriptor increments next to the MMAs
__device__ __forceinline__ void mma_i8_x4(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t astep, uint32_t bstep, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q, t;\n\t.reg .b64 a1, b1, a2, b2, a3, b3, sa, sb;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "setp.eq.b32 t, 0, 0;\n\t"
      "cvt.u64.u32 sa, %3;\n\tcvt.u64.u32 sb, %4;\n\t"
      "add.s64 a1, %1, sa;\n\tadd.s64 b1, %2, sb;\n\t"
      "add.s64 a2, a1, sa;\n\tadd.s64 b2, b1, sb;\n\t"
      "add.s64 a3, a2, sa;\n\tadd.s64 b3, b2, sb;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %5, p;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::i8 [%0], a1, b1, %5, t;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::i8 [%0], a2, b2, %5, t;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::i8 [%0], a3, b3, %5, t;\n\t}\n"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(astep), "r"(bstep), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit_elect(uint32_t bar) {
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, int (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__host__ __device__ inline size_t i8_q_tile_bytes(int NS, int KA) { return (size_t)NS * 4 * (2 * KA) * 128 + CN * sizeof(double); }
__host__ __device__ inline size_t i8_smem_bytes(int NS, int DC, int KA, int NB) {
  return (size_t)NB * 2048 * DC + ((size_t)NS * 4 * (2 * KA) * 128) + (size_t)NS * 4096 + 2 * CN * sizeof(double) /* scales, two buffers */ +
         32 * sizeof(double) /* exp table */ + 4 * CN * sizeof(double) /* lp partials */ + (size_t)(2 * NB + 8) * 8 + 16;
}

// ---- the main kernel ------------------------------------------------------------------------------------------------
template <int NS, bool TRACE>
__global__ void __launch_bounds__(NTH, 1) glm_i8_kernel(const GlmI8Args a) {
  auto stamp = [&](int role, uint32_t it, int k) {
    if (TRACE && blockIdx.x == 0 && it < (uint32_t)kTraceIt) a.trace[((size_t)role * kTraceIt + it) * 8 + k] = clock64();
  };
  extern __shared__ __align__(128) uint8_t smem_i8[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int DC = a.DC, KA = a.KA, DCQ = 2 * KA, NB = a.NB;
  const uint32_t slice_bytes = 2048u * DC;
  const uint32_t q_digit_bytes = (uint32_t)NS * 4 * DCQ * 128;
  uint8_t* ring = smem_i8;
  uint8_t* qs = ring + (size_t)NB * slice_bytes;
  uint8_t* dms = qs + q_digit_bytes;
  double* scs = reinterpret_cast<double*>(dms + NS * 4096);   // [2][CN]
  double* exptab = scs + 2 * CN;
  double* red = exptab + 32;                                   // [4][CN]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(red + 4 * CN);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * NB + 8);
  const uint32_t bar0 = smem_u32(bars);
  auto x_full = [&](int s) { return bar0 + 8u * s; };
  auto x_empty = [&](int s) { return bar0 + 8u * (NB + s); };
  const uint32_t q_full = bar0 + 8u * (2 * NB), q_empty = q_full + 8, a_full = q_full + 16, a_empty = q_full + 24, dm_full = q_full + 32,
                 dm_empty = q_full + 40, g_full = q_full + 48, g_empty = q_full + 56;

  const int p = blockIdx.x % a.P, kq = blockIdx.x / a.P;
  const int n_list = a.n_list ? *a.n_list : a.n_fixed;
  const int n_ct = (n_list + CN - 1) / CN;
  const int n_my = n_ct > kq ? (n_ct - kq + KQ - 1) / KQ : 0;
  const int t_begin = (int)((long long)p * a.T / a.P), t_end = (int)((long long)(p + 1) * a.T / a.P);
  const int n_rt = t_end - t_begin;
  const uint32_t n_it = (uint32_t)n_my * (uint32_t)n_rt;
  if (n_it == 0) return;

  if (tid == 0) {
    for (int s = 0; s < 2 * NB; s++) mbar_init(bar0 + 8u * s, 1);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    mbar_init(a_full, 1);
    mbar_init(a_empty, NEW);
    mbar_init(dm_full, NEW);
    mbar_init(dm_empty, 1);
    mbar_init(g_full, 1);
    mbar_init(g_empty, NEW);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_ptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  exp_table_load(exptab, tid);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_ptr;

  if (warp == 0) {
    // ===== producer: q planes of each chain tile, X planes of each row tile through the ring =====
    // (one thread; every index is carried incrementally: a division or a modulo costs this lone thread more than an MMA takes)
    if (lane == 0) {
      const size_t q_tile = i8_q_tile_bytes(NS, KA);
      const uint32_t ring_u = smem_u32(ring);
      uint32_t slot = 0, pass = 0;   // pass over the ring: use number of `slot`
      // every CTA runs the same schedule, so all 148 would ask L2 for their six planes (86 KB each) at the same moment of every tile
      // and a plane would take ~4 500 cycles to arrive (r02t / r03c); offsets spread over one tile period keep the bursts apart
      __nanosleep((blockIdx.x % 16u) * 200u);
      const uint64_t pol = l2_policy_evict_last();
      const uint8_t* qsrc = a.Qimg + (size_t)kq * q_tile;
      for (int j = 0; j < n_my; j++, qsrc += (size_t)KQ * q_tile) {
        if (j > 0) mbar_wait_relaxed(q_empty, (uint32_t)(j - 1) & 1u);
        mbar_expect_tx(q_full, q_digit_bytes + CN * (unsigned)sizeof(double));
        tma_load_1d(smem_u32(qs), qsrc, q_digit_bytes, q_full);
        tma_load_1d(smem_u32(scs + (j & 1) * CN), qsrc + q_digit_bytes, CN * (unsigned)sizeof(double), q_full);
        const uint8_t* xsrc = a.Ximg + (size_t)t_begin * NS * slice_bytes;
        for (int rt = 0; rt < n_rt; rt++) {
#pragma unroll
          for (int i = 0; i < NS; i++, xsrc += slice_bytes) {
            if (pass > 0) mbar_wait_relaxed(x_empty(slot), (pass - 1) & 1u);
            stamp(2, (uint32_t)(j * n_rt + rt), i);
            mbar_expect_tx(x_full(slot), slice_bytes);
            tma_load_1d_hint(ring_u + slot * slice_bytes, xsrc, slice_bytes, x_full(slot), pol);
            if (++slot == (uint32_t)NB) {
              slot = 0;
              pass++;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (warp 1, converged; one elected lane issues): first product of tile it + 1 before the second product of tile it =====
    {
      const uint32_t ring_u = smem_u32(ring), qs_u = smem_u32(qs), dm_u = smem_u32(dms);
      const uint64_t bdesc_a = make_desc(qs_u, 128, DCQ * 128), bdesc_b = make_desc(dm_u, 128, 2048);
      const uint64_t adesc_a0 = make_desc(ring_u, 128, DC * 128), adesc_b0 = make_desc(ring_u, DC * 128, 128);
      const uint32_t slice16 = slice_bytes >> 4, kstep_b = (uint32_t)(4 * DC * 8);
      // cursors of the two issue streams (both walk the ring in the same order)
      uint32_t a_slot = 0, a_par = 0, a_it = 0, b_slot = 0, b_it = 0;
      int a_rt = 0, a_j = 0, b_rt = 0, b_j = 0;
      auto issue_a = [&]() {
        stamp(0, a_it, 0);
        if (a_it > 0) mbar_wait_warp(a_empty, (a_it - 1) & 1u);
        if (a_rt == 0) mbar_wait_warp(q_full, (uint32_t)a_j & 1u);
        tc_fence_after();
        stamp(0, a_it, 1);
#pragma unroll
        for (int i = 0; i < NS; i++) {
          mbar_wait_warp(x_full(a_slot), a_par);
          if (i == 0) stamp(0, a_it, 2);
          tc_fence_after();
          const uint64_t adesc = adesc_a0 + (uint64_t)(a_slot * slice16);
          if (KA == 4) {
            mma_i8_x4(tbase + ACOL + 32 * i, adesc, bdesc_a, 16u, 16u, idesc_i8(128, 32 * (NS - i), 0), i > 0 ? 1u : 0u);
          } else {
            for (int k = 0; k < KA; k++)
              mma_i8(tbase + ACOL + 32 * i, adesc + (uint64_t)(k * 16), bdesc_a + (uint64_t)(k * 16), idesc_i8(128, 32 * (NS - i), 0), (i > 0 || k > 0) ? 1u : 0u);
          }
          if (++a_slot == (uint32_t)NB) {
            a_slot = 0;
            a_par ^= 1u;
          }
        }
        tc_commit_elect(a_full);
        stamp(0, a_it, 3);
        a_it++;
        if (++a_rt == n_rt) {
          tc_commit_elect(q_empty);
          a_rt = 0;
          a_j++;
        }
      };
      auto issue_b = [&]() {
        stamp(0, b_it, 4);
        mbar_wait_warp(dm_full, b_it & 1u);
        stamp(0, b_it, 5);
        if (b_rt == 0 && b_j > 0) mbar_wait_warp(g_empty, (uint32_t)(b_j - 1) & 1u);
        tc_fence_after();
#pragma unroll
        for (int i = 0; i < NS; i++) {
          const uint64_t adesc = adesc_b0 + (uint64_t)(b_slot * slice16);
          mma_i8_x4(tbase + GCOL + 32 * i, adesc, bdesc_b, kstep_b, 32u, idesc_i8(128, 32 * (NS - i), 1), (b_rt > 0 || i > 0) ? 1u : 0u);
          tc_commit_elect(x_empty(b_slot));
          if (++b_slot == (uint32_t)NB) b_slot = 0;
        }
        tc_commit_elect(dm_empty);
        stamp(0, b_it, 6);
        b_it++;
        if (++b_rt == n_rt) {
          tc_commit_elect(g_full);
          b_rt = 0;
          b_j++;
        }
      };
      issue_a();
      for (uint32_t it = 0; it < n_it; it++) {
        if (it + 1 < n_it) issue_a();
        issue_b();
      }
    }
  } else {
    // ===== epilogue: thread = row r of the tile (TMEM lane), chains [8 wg, 8 wg + 8) of the chain tile =====
    const int ew = warp - 2, wg = ew >> 2, quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int et = ew * 32 + lane;
    const uint32_t lane_base = tbase + ((uint32_t)(quarter * 32) << 16);
    const uint32_t dm_row = smem_u32(dms) + (uint32_t)(wg >> 1) * 2048u + (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u + (uint32_t)(wg & 1) * 8u;
    const double gscale = ldexp(1.0, I8Fmt<NS>::SEXP) * (r < a.Dk ? a.colscale[r] : 0.0);
    double lprod[ECH], lsum[ECH];
    uint32_t it = 0;
    int y_next = (long long)t_begin * TRW + r < a.N ? a.y[(long long)t_begin * TRW + r] : 1;
    for (int j = 0; j < n_my; j++) {
      const int slot0 = (kq + KQ * j) * CN;
#pragma unroll
      for (int c = 0; c < ECH; c++) {
        lprod[c] = 1.0;
        lsum[c] = 0.0;
      }
      int n_pad = 0;
      // (no wait on q_full here: with one row tile per chain tile the MMA thread may run two chain tiles ahead of a warp that is
      // still in its flush, and a parity wait cannot tell phase j from phase j + 2 -- r02u deadlock.  The scales of chain tile j
      // are in place before the first product of the tile is issued, and a_full is waited for before they are read)
      const double* sc = scs + (j & 1) * CN + wg * ECH;
      for (int rt = 0; rt < n_rt; rt++, it++) {
        const long long row = (long long)(t_begin + rt) * TRW + r;
        const bool valid = row < a.N;
        const int negmask = (valid && y_next == 0) ? (int)0x80000000 : 0;
        n_pad += valid ? 0 : 1;
        {   // the observation of the next tile's row: its latency hides behind this tile's work
          const long long rown = rt + 1 < n_rt ? row + TRW : (long long)t_begin * TRW + r;
          y_next = rown < a.N ? a.y[rown] : 1;
        }
        if (et == 0 || et == 96) stamp(et ? 3 : 1, it, 0);
        mbar_wait(a_full, it & 1u);
        tc_fence_after();
        if (et == 0 || et == 96) stamp(et ? 3 : 1, it, 1);
        long long Sv[ECH];   // the recombination is integer-only, so that it can run under the previous tile's second product
        {
          int v[NS][ECH];
#pragma unroll
          for (int k = 0; k < NS; k++) tmem_ld8(lane_base + ACOL + 32 * k + ECH * wg, v[k]);
          tmem_wait_ld();
          tc_fence_before();
          if (I8_EARLY_RELEASE) {
            __syncwarp();
            if (lane == 0) mbar_arrive(a_empty);
          }
          if (et == 0 || et == 96) stamp(et ? 3 : 1, it, 2);
#pragma unroll
          for (int c = 0; c < ECH; c++) {
            int d[NS];
#pragma unroll
            for (int k = 0; k < NS; k++) d[k] = v[k][c];
            Sv[c] = i8_combine<NS>(d);
          }
        }
        if (et == 0 || et == 96) stamp(et ? 3 : 1, it, 6);
        // ---- the FP64 phase: conversion + link of all eight chains, four at a time.  Scalar FP64 is starved while the tensor core
        // runs (tools/microbench/fp64_link.cu: 72x): what precedes (read, integer recombination) and what follows (digit planes,
        // transposes, stores, hand-over) is integer-only and costs nothing extra under an MMA stream; see I8_EARLY_RELEASE.
        double dv[ECH];
#pragma unroll
        for (int g = 0; g < ECH / 4; g++) {
          double e4[4], d4[4], pr[4], sm[4];
#pragma unroll
          for (int c = 0; c < 4; c++) {
            e4[c] = (double)Sv[4 * g + c] * sc[4 * g + c];   // (I2F.F64.S64 + DMUL: the magic-number conversion costs two more FP64
                                                             // instructions and measured slower, r02z)
            pr[c] = lprod[4 * g + c];
            sm[c] = lsum[4 * g + c];
          }
          logit_terms4(pr, sm, e4, negmask, exptab, d4);
          if (g == 0 && (et == 0 || et == 96)) stamp(et ? 3 : 1, it, 7);
#pragma unroll
          for (int c = 0; c < 4; c++) {
            lprod[4 * g + c] = pr[c];
            lsum[4 * g + c] = sm[c];
            dv[4 * g + c] = d4[c];
          }
        }
        if (!I8_EARLY_RELEASE) {
          __syncwarp();
          if (lane == 0) mbar_arrive(a_empty);   // the next tile's first product may start (and overwrite the Eta diagonals)
        }
        if (et == 0 || et == 96) stamp(et ? 3 : 1, it, 3);
        if (it > 0) mbar_wait(dm_empty, (it - 1) & 1u);   // the second product of the previous tile has read the dm planes
        if (et == 0 || et == 96) stamp(et ? 3 : 1, it, 4);
#pragma unroll
        for (int g = 0; g < ECH / 4; g++) {
          double d4[4];
#pragma unroll
          for (int c = 0; c < 4; c++) d4[c] = dv[4 * g + c];
          uint32_t pk[NS];
          i8_digit_planes4<NS>(d4, pk);
#pragma unroll
          for (int s = 0; s < NS; s++)
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(dm_row + (uint32_t)s * 4096u + (uint32_t)g * 4u), "r"(pk[s]) : "memory");
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(dm_full);
        if (et == 0 || et == 96) stamp(et ? 3 : 1, it, 5);
      }
      // ---- end of the chain tile: partial gradient and log-likelihood of (row group p, chain tile)
      mbar_wait(g_full, (uint32_t)j & 1u);
      tc_fence_after();
      {
        int v[NS][ECH];
#pragma unroll
        for (int k = 0; k < NS; k++) tmem_ld8(lane_base + GCOL + 32 * k + ECH * wg, v[k]);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(g_empty);
#pragma unroll
        for (int c = 0; c < ECH; c++) {
          // sum_k diag_k 256^(NS-1-k): the top diagonal separately in FP64 -- with up to 21 k rows per group it reaches 2^30, and
          // 2^30 2^40 does not fit 64 bits; the rest stays below 2^61
          long long S = v[1][c];
#pragma unroll
          for (int k = 2; k < NS; k++) S = S * 256 + v[k][c];
          const int slot = slot0 + ECH * wg + c;
          if (r < a.Dk && slot < n_list)
            a.Gp[((size_t)p * a.cap + slot) * a.Dk + r] = fma((double)v[0][c], (double)(1ll << (8 * (NS - 1))), (double)S) * gscale;
        }
      }
#pragma unroll
      for (int c = 0; c < ECH; c++) {
        LogitAcc la;
        la.prod = lprod[c];
        la.sum = lsum[c];
        la.expo = 0;
        logit_acc_fold(la);
        double v = logit_acc_value(la, n_pad);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if (lane == 0) red[quarter * CN + ECH * wg + c] = v;
      }
      asm volatile("bar.sync 1, 512;" ::: "memory");
      if (et < CN && slot0 + et < n_list) a.LPp[(size_t)p * a.cap + slot0 + et] = ((red[et] + red[CN + et]) + red[2 * CN + et]) + red[3 * CN + et];
      asm volatile("bar.sync 1, 512;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
}

// ---- per launch: q of the requesting chains -> digit planes (K-major B operand of the first product) + scales ------------
template <int NS>
__global__ void __launch_bounds__(128) glm_i8_qprep_kernel(const double* __restrict__ Q, int ldq, int D, const int* __restrict__ list,
                                                           const int* __restrict__ n_list_p, int n_fixed, const double* __restrict__ colscale,
                                                           int KA, uint8_t* __restrict__ Qimg) {
  const int n_list = n_list_p ? *n_list_p : n_fixed;
  const int ct = blockIdx.x;
  if (ct * CN >= n_list) return;
  const int DCQ = 2 * KA, KP = DCQ * 16;
  const size_t q_tile = i8_q_tile_bytes(NS, KA);
  uint8_t* img = Qimg + (size_t)ct * q_tile;
  double* scs = reinterpret_cast<double*>(img + (size_t)NS * 4 * DCQ * 128);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = warp; c < CN; c += 4) {
    const int slot = ct * CN + c;
    const bool live = slot < n_list;
    const double* q = live ? Q + (size_t)(list ? list[slot] : slot) * ldq : nullptr;
    double v[4];
    unsigned long long mx = 0;
    bool finite = true;
#pragma unroll
    for (int h = 0; h < 4; h++) {
      const int k = lane + 32 * h;
      v[h] = (live && k < D) ? q[k] * colscale[k] : 0.0;
      const unsigned long long bits = (unsigned long long)__double_as_longlong(v[h]) & 0x7fffffffffffffffull;
      finite = finite && bits < 0x7ff0000000000000ull;
      mx = bits > mx ? bits : mx;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, mx, o);
      mx = other > mx ? other : mx;
    }
    finite = __all_sync(0xffffffffu, finite);
    int eq = (int)(mx >> 52) - 1022;   // |v| < 2^eq
    if (eq < -900) eq = -900;
    if (lane == 0) scs[c] = finite ? ldexp(1.0, eq + I8Fmt<NS>::SEXP) : __longlong_as_double(0x7ff8000000000000ll);
#pragma unroll
    for (int h = 0; h < 4; h++) {
      const int k = lane + 32 * h;
      if (k >= KP) continue;
      long long V = finite ? __double2ll_rn(ldexp(v[h], I8Fmt<NS>::F - eq)) : 0;
      int dg[NS];
#pragma unroll
      for (int s = NS - 1; s > 0; s--) {
        const int d = (int)(((V + 128) & 255) - 128);
        dg[s] = d;
        V = (V - d) >> 8;
      }
      dg[0] = (int)V;
#pragma unroll
      for (int s = 0; s < NS; s++)
        img[(((size_t)s * 4 + (c >> 3)) * DCQ + (k >> 4)) * 128 + (c & 7) * 16 + (k & 15)] = (uint8_t)(int8_t)dg[s];
    }
  }
}

// ---- once per data set: column scales and the digit-plane image of X ----------------------------------------------------
__global__ void glm_i8_colmax_kernel(const double* __restrict__ X, long long N, int ldx, int Dk, unsigned long long* __restrict__ colmax) {
  const int k = threadIdx.x;   // blockDim.x == ldx-ish: one thread per column, rows strided over the grid
  if (k >= Dk) return;
  unsigned long long mx = 0;
  for (long long row = blockIdx.x; row < N; row += gridDim.x) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(X[(size_t)row * ldx + k]) & 0x7fffffffffffffffull;
    mx = bits > mx ? bits : mx;
  }
  atomicMax(colmax + k, mx);
}
__global__ void glm_i8_colscale_kernel(const unsigned long long* __restrict__ colmax, int Dk, double* __restrict__ colscale, int* __restrict__ bad) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= Dk) return;
  const unsigned long long mx = colmax[k];
  if (mx >= 0x7ff0000000000000ull) *bad = 1;   // non-finite design entry: the INT8 image cannot represent it
  int e = (int)(mx >> 52) - 1022;
  if (mx == 0 || e < -900) e = -900;
  colscale[k] = ldexp(1.0, e);
}
template <int NS>
__global__ void glm_i8_ximg_kernel(const double* __restrict__ X, long long N, int ldx, int Dk, int DC, int T, const double* __restrict__ colscale,
                                   uint8_t* __restrict__ Ximg) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (padded row, 16-column chunk)
  if (idx >= (long long)T * TRW * DC) return;
  const long long row = idx / DC;
  const int dc = (int)(idx - row * DC);
  const int t = (int)(row / TRW), rr = (int)(row - (long long)t * TRW);
  uint8_t out[NS][16];
#pragma unroll
  for (int e = 0; e < 16; e++) {
    const int k = dc * 16 + e;
    long long V = 0;
    if (row < N && k < Dk) V = __double2ll_rn(ldexp(X[(size_t)row * ldx + k] / colscale[k], I8Fmt<NS>::F));
#pragma unroll
    for (int s = NS - 1; s > 0; s--) {
      const int d = (int)(((V + 128) & 255) - 128);
      out[s][e] = (uint8_t)(int8_t)d;
      V = (V - d) >> 8;
    }
    out[0][e] = (uint8_t)(int8_t)V;
  }
  const size_t slice_bytes = (size_t)2048 * DC;
#pragma unroll
  for (int s = 0; s < NS; s++) {
    uint8_t* dst = Ximg + ((size_t)t * NS + s) * slice_bytes + ((size_t)(rr >> 3) * DC + dc) * 128 + (rr & 7) * 16;
    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(out[s]);
  }
}

}  // namespace

struct GlmI8State {
  int NS = 6, N = 0, Dk = 0, DC = 0, KA = 0, T = 0, P = 0, cap = 0, NB = 0;
  uint8_t* Ximg = nullptr;
  uint8_t* Qimg = nullptr;
  double* colscale = nullptr;
  size_t smem = 0;
};

void glm_i8_destroy(void* st) {
  GlmI8State* s = static_cast<GlmI8State*>(st);
  if (!s) return;
  cudaFree(s->Ximg);
  cudaFree(s->Qimg);
  cudaFree(s->colscale);
  delete s;
}

int glm_i8_row_groups(const void* st) { return static_cast<const GlmI8State*>(st)->P; }

// nullptr when the formulation does not apply (other links, more than 128 design columns, row groups too long for INT32 sums,
// non-finite design entries): the caller keeps the FP64 tensor kernel
void* glm_i8_create(const ModelData& m, int Dk, int cap, int sm_count, int max_row_groups, cudaStream_t stream) {
  if (m.kind != MODEL_LOGISTIC || Dk > 128 || m.N < 1) return nullptr;
  const char* nsv = getenv("B200NUTS_I8_SLICES");
  const int NS = nsv && atoi(nsv) == 5 ? 5 : 6;
  GlmI8State* s = new GlmI8State;
  s->NS = NS;
  s->N = m.N;
  s->Dk = Dk;
  s->DC = (Dk + 15) / 16;
  s->KA = (Dk + 31) / 32;
  s->T = (m.N + TRW - 1) / TRW;
  int P = sm_count / KQ;
  if (P < 1) P = 1;
  if (P > s->T) P = s->T;
  if (P > max_row_groups) P = max_row_groups;
  s->P = P;
  s->cap = (cap + CN - 1) / CN * CN;
  const long long rows_per_group = ((long long)(s->T + P - 1) / P) * TRW;
  // INT32 sums: |digit product| <= 2^14, NS pairs on the last diagonal; log-likelihood: at most ~900 factors in (1, 2] per running product
  if (rows_per_group * 16384ll * NS >= (1ll << 31) || rows_per_group / TRW > 900) {
    delete s;
    return nullptr;
  }
  const size_t limit = 232448;
  int NB = 2 * NS;
  while (NB < 4 * NS && i8_smem_bytes(NS, s->DC, s->KA, NB + 1) <= limit) NB++;
  if (i8_smem_bytes(NS, s->DC, s->KA, NB) > limit) {
    delete s;
    return nullptr;
  }
  s->NB = NB;
  s->smem = i8_smem_bytes(NS, s->DC, s->KA, NB);
  // (a persisting-L2 set-aside for the digit planes, cudaLimitPersistingL2CacheSize, measured within noise -- r03l -- and would take
  // L2 capacity from every other kernel of the process: not used)
  const size_t slice_bytes = (size_t)2048 * s->DC;
  unsigned long long* colmax = nullptr;
  int* bad = nullptr;
  bool ok = cudaMalloc(&s->Ximg, (size_t)s->T * NS * slice_bytes + 4096) == cudaSuccess &&
            cudaMalloc(&s->Qimg, (size_t)(s->cap / CN) * i8_q_tile_bytes(NS, s->KA)) == cudaSuccess &&
            cudaMalloc(&s->colscale, sizeof(double) * Dk) == cudaSuccess && cudaMalloc(&colmax, sizeof(unsigned long long) * Dk + sizeof(int)) == cudaSuccess;
  if (ok) {
    bad = reinterpret_cast<int*>(colmax + Dk);
    cudaMemsetAsync(colmax, 0, sizeof(unsigned long long) * Dk + sizeof(int), stream);
    glm_i8_colmax_kernel<<<592, 128, 0, stream>>>(m.X, m.N, m.ldx, Dk, colmax);
    glm_i8_colscale_kernel<<<1, 128, 0, stream>>>(colmax, Dk, s->colscale, bad);
    const long long n = (long long)s->T * TRW * s->DC;
    if (NS == 6) glm_i8_ximg_kernel<6><<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(m.X, m.N, m.ldx, Dk, s->DC, s->T, s->colscale, s->Ximg);
    else glm_i8_ximg_kernel<5><<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(m.X, m.N, m.ldx, Dk, s->DC, s->T, s->colscale, s->Ximg);
    int h_bad = 0;
    ok = cudaMemcpyAsync(&h_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, stream) == cudaSuccess && cudaStreamSynchronize(stream) == cudaSuccess && !h_bad;
  }
  cudaFree(colmax);
  if (!ok) {
    cudaGetLastError();
    glm_i8_destroy(s);
    return nullptr;
  }
  return s;
}

// Gp / LPp partials of the requesting chains for the row groups [0, glm_i8_row_groups): the caller reduces and finishes as for the
// other kernels.  Returns the number of kernels launched.
int launch_glm_i8(const ModelData& m, const GradBatch& b, const GlmWorkspace& ws, cudaStream_t stream) {
  const GlmI8State* s = static_cast<const GlmI8State*>(ws.i8);
  GlmI8Args a;
  a.Ximg = s->Ximg;
  a.Qimg = s->Qimg;
  a.colscale = s->colscale;
  a.y = m.yi;
  a.n_list = b.n_list;
  a.Gp = ws.Gp;
  a.LPp = ws.LPp;
  a.n_fixed = b.C;
  a.N = m.N;
  a.Dk = s->Dk;
  a.DC = s->DC;
  a.KA = s->KA;
  a.cap = ws.C;
  a.P = s->P;
  a.T = s->T;
  a.NB = s->NB;
  const int n_tiles = (b.C + CN - 1) / CN;
  a.trace = nullptr;
  // developer aid: B200NUTS_I8_TRACE=<file> writes the clock stamps of CTA 0 of the B200NUTS_I8_TRACE_LAUNCH-th launch (default 3)
  static const char* trace_path = getenv("B200NUTS_I8_TRACE");
  static int trace_countdown = getenv("B200NUTS_I8_TRACE_LAUNCH") ? atoi(getenv("B200NUTS_I8_TRACE_LAUNCH")) : 3;
  const bool trace = trace_path && s->NS == 6 && trace_countdown-- == 0;
  if (trace) {
    cudaMalloc(&a.trace, sizeof(long long) * 4 * kTraceIt * 8);
    cudaMemsetAsync(a.trace, 0, sizeof(long long) * 4 * kTraceIt * 8, stream);
  }
  if (s->NS == 6) {
    glm_i8_qprep_kernel<6><<<n_tiles, 128, 0, stream>>>(b.Q, b.ld, m.D, b.list, b.n_list, b.C, s->colscale, s->KA, s->Qimg);
    auto k = trace ? glm_i8_kernel<6, true> : glm_i8_kernel<6, false>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem);
    k<<<s->P * KQ, NTH, s->smem, stream>>>(a);
  } else {
    glm_i8_qprep_kernel<5><<<n_tiles, 128, 0, stream>>>(b.Q, b.ld, m.D, b.list, b.n_list, b.C, s->colscale, s->KA, s->Qimg);
    auto k = glm_i8_kernel<5, false>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem);
    k<<<s->P * KQ, NTH, s->smem, stream>>>(a);
  }
  if (trace) {
    static long long host[4 * kTraceIt * 8];
    cudaStreamSynchronize(stream);
    cudaMemcpy(host, a.trace, sizeof(host), cudaMemcpyDeviceToHost);
    cudaFree(a.trace);
    if (FILE* f = fopen(trace_path, "wb")) {
      fwrite(host, 1, sizeof(host), f);
      fclose(f);
    }
  }
  return 2;
}

}  // namespace b200
