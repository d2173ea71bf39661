// Hierarchical IRT 2PL log-density + gradient for the lock-step engine (config c4).
//
//   parameters (declaration order = unconstrained layout, positive ones as logs with Jacobian +u, lb_constrain.hpp:97):
//     sigma_theta, theta[I], sigma_a, a[J], mu_b, sigma_b, b[J]                              D = 4 + I + 2J
//   model: sigma_* ~ cauchy(0,2) (cauchy_lpdf.hpp:80-86); theta ~ normal(0, sigma_theta), b ~ normal(mu_b, sigma_b)
//          (normal_lpdf.hpp:83-89, keeps -N log sigma); a ~ lognormal(0, sigma_a) (lognormal_lpdf.hpp:72-80, keeps
//          N * NEG_LOG_SQRT_TWO_PI even under propto); mu_b ~ normal(0,5);
//          y[i,j] ~ bernoulli_logit(a_j (theta_i - b_j)) (bernoulli_logit_lpmf.hpp:73-89, cutoff +-20).
//
// Work distribution inside a CTA: 256 threads = 32 groups of 8 lanes; a group takes persons i = g, g+32, ... of a person block;
// inside a group lane s owns items j = s, s+8, ... (IPL = ceil(J/8) items per lane: J = 100 -> 13, 96 % of the lane slots busy),
// whose a/b gradients accumulate in registers; the theta_i gradient is a 3-step shuffle reduction over the group.  The
// per-response link is the product-form one of fastmath.cuh (21 FP64 + ~25 other instructions: on sm_100 an FP64 instruction
// holds its pipe for two issue slots, so the two are balanced).  y is shared by all chains (int8, L2-resident).
//
// Two kernels, one arithmetic (irt_body.cuh):
//   irt_grad_kernel   THROUGHPUT form: one CTA per requesting chain, launched once per lock-step over the compacted list.
//   irt_solo_kernel   LATENCY form for the stragglers of a run: in the reference every chain is its own task and never waits for
//                     another (hmc_nuts_diag_e_adapt.hpp:387-401); in a lock-step engine the number of steps of a run is the
//                     leapfrog count of its SLOWEST chain (r01h: 30 690 steps for 30 transitions of 8192 chains, mean 31
//                     leapfrogs per transition), so once only a few chains are left each of them gets a CLUSTER of eight
//                     CTAs (eight SMs) that runs it to the end of the call on the device: CTA b evaluates person block b,
//                     CTA 0 adds the block partials through distributed shared memory in the canonical order and advances the
//                     chain with all its 256 threads (solo_leaf.cuh); no launches, no host round trips, no other chain to
//                     wait for.  Bit-identical to the throughput form + advance_kernel by construction.
#define B200_VEC_UNROLL 8   // the generic advance() of the latency kernel must round like advance_big.cu's (it does: unrolling
#define B200_WITH_DENSE 0   // never reorders a sum) and is built without the dense / diagnostics / static-HMC branches
#include <cstdlib>

#ifndef B200_SIMT_EMUL
#include <cooperative_groups.h>
#endif

#include "chains.cuh"
#include "irt_body.cuh"
#include "solo_leaf.cuh"

namespace b200 {

int irt_max_items() { return IRT_MAXJ; }

// OCC = 1: item parameters in registers, one CTA per SM.  OCC = 2 / 3: exp(a), b in shared memory and a 128- / 80-register budget,
// two / three CTAs (16 / 24 warps) per SM: the one-CTA form is latency-bound at two warps per sub-partition (r01c / r01d captures).
template <int IPL, int OCC>
__global__ void __launch_bounds__(IRT_NT, OCC) irt_grad_kernel(const unsigned short* __restrict__ y, int I, int J, int D,
                                                             const double* __restrict__ Q, double* __restrict__ G,
                                                             double* __restrict__ zV, const int* __restrict__ active,
                                                             const int* __restrict__ list, const int* __restrict__ n_list, int ld) {
  __shared__ IrtShared sm;
  int c = blockIdx.x;
  if (list) {
    if (c >= *n_list) return;
    c = list[c];
  } else if (active && !active[c]) {
    return;
  }
  irt_chain_body<IPL, OCC>(sm, y, I, J, D, Q + (size_t)c * ld, G + (size_t)c * ld, zV + c);
}

int launch_irt_grad(const ModelData& m, const GradBatch& b, cudaStream_t stream) {
  const int ipl = (m.K + IRT_G - 1) / IRT_G;
  // B200NUTS_IRT_VARIANT: 1 / 2 / 3 CTAs per SM; default "auto": three per SM when the launch is known to be dense (24 resident
  // warps hide the link's dependency chains: r02i, 8192 chains: gradient stage 2.51 s with three, 2.74 s with two, 3.68 s with one),
  // one per SM when only a few chains are left (lower latency of a single CTA).  The forms are the same arithmetic in the same
  // order (irt_body.cuh; GPU test), so the choice never shows in a chain's draws.  Read per launch (a handful of ns next to a
  // > 100 us kernel): tests switch it.
  const char* var = getenv("B200NUTS_IRT_VARIANT");
  int occ = (var && var[0] != 'a') ? atoi(var) : ((b.hint_active < 0 ? b.C : b.hint_active) >= 296 ? 3 : 1);
#define IRT_LAUNCH(IPL)                                                                                                          \
  {                                                                                                                              \
    auto k1 = irt_grad_kernel<IPL, 1>;                                                                                           \
    auto k2 = irt_grad_kernel<IPL, 2>;                                                                                           \
    auto k3 = irt_grad_kernel<IPL, 3>;                                                                                           \
    if (occ == 3) IRT_KERNEL_LAUNCH(k3, b.C, stream, m.ypk, m.N, m.K, m.D, b.Q, b.G, b.zV, b.active, b.list, b.n_list, b.ld);     \
    else if (occ == 2) IRT_KERNEL_LAUNCH(k2, b.C, stream, m.ypk, m.N, m.K, m.D, b.Q, b.G, b.zV, b.active, b.list, b.n_list, b.ld); \
    else IRT_KERNEL_LAUNCH(k1, b.C, stream, m.ypk, m.N, m.K, m.D, b.Q, b.G, b.zV, b.active, b.list, b.n_list, b.ld);              \
  }
  if (ipl <= 2) IRT_LAUNCH(2)
  else if (ipl <= 4) IRT_LAUNCH(4)
  else if (ipl <= 7) IRT_LAUNCH(7)
  else if (ipl <= 10) IRT_LAUNCH(10)
  else if (ipl <= 13) IRT_LAUNCH(13)
  else IRT_LAUNCH(16)
#undef IRT_LAUNCH
  return 1;
}

// =====================================================================================================================
// latency form
// =====================================================================================================================
// one out-of-line copy of the generic state machine for all instantiations of the latency kernel (it runs a few times per
// transition; the leaf path is solo_leaf_step)
static __device__ __noinline__ int advance_out_of_line(ChainView v, const SamplerCfg* cfg, int lane) { return advance(v, *cfg, lane); }

constexpr int IRT_SOLO_MAX_PB = 512;   // persons per block the theta-gradient staging holds (I <= 4096)

struct SoloShared {
  IrtShared irt;
  double th[IRT_SOLO_MAX_PB];   // cluster form: this CTA's theta gradients, collected by CTA 0
  ChainS chain;                 // CTA 0: the chain's scalar block for the duration of the launch
  SoloFlags fl;
  int more;
};

// CL = 8: one cluster of eight CTAs per chain (hardware only).  CL = 1: one CTA per chain walks the eight person blocks itself
// (same bits; the CPU test rig runs this one, and it is the fallback where a cluster launch is not available).
template <int IPL, int CL>
__global__ void __launch_bounds__(IRT_NT, 1) irt_solo_kernel(SamplerCfg cfg, DeviceChains d, const unsigned short* __restrict__ y, int I, int J,
                                                             const int* __restrict__ list, const int* __restrict__ n_list) {
  __shared__ __align__(16) SoloShared sh;
  B200_DYN_SMEM(red);   // CTA 0: 6 * ld doubles (solo_leaf_step)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, sub = tid & (IRT_G - 1);
  const int rank = CL == 1 ? 0 : (int)(blockIdx.x % CL);
  const int slot_i = (int)(blockIdx.x / CL);
  if (slot_i >= *n_list) return;   // the whole cluster leaves together
  const int c = list[slot_i];
#ifndef B200_SIMT_EMUL
  cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
#define SOLO_CLUSTER_SYNC() do { if (CL > 1) cluster.sync(); else __syncthreads(); } while (0)
#else
#define SOLO_CLUSTER_SYNC() __syncthreads()
#endif
  ChainView v = make_view(d, cfg, c);
  ChainS* gs = v.s;
  if (rank == 0 && warp == 0) warp_copy_words(&sh.chain, gs, kChainWords, lane);
  __syncthreads();
  v.s = &sh.chain;
#ifdef B200_SIMT_EMUL  // the rig's lanes are free-running: warp 0's lanes each work on a replica of the scalar block inside advance()
  ChainS simt_replica;
#endif
  const int PB = (I + IRT_NB - 1) / IRT_NB;
  for (;;) {
    // ---- gradient at z.q (written by the advance step, possibly earlier in this very kernel: coherent loads)
    {
      const IrtChainPtrs cp = irt_chain_ptrs<true>(v.q, I, J);
      double a[IPL], b[IPL];
      irt_prologue<IPL, 1, true>(sh.irt, cp, J, a, b);
      if (CL == 1) {
        IrtWarpTotals tot;
        for (int blk = 0; blk < IRT_NB; blk++) {
          double va[IPL], vb[IPL], w_prod, w_sum, w_s2;
          int w_expo;
          irt_block<IPL, 1, true>(sh.irt, cp, y, I, J, blk, a, b, v.g, va, vb, w_prod, w_expo, w_sum, w_s2);
          if (lane < IRT_G) {
#pragma unroll
            for (int t = 0; t < IPL; t++) {
              sh.irt.ga[warp][sub + IRT_G * t] += va[t];
              sh.irt.gb[warp][sub + IRT_G * t] += vb[t];
            }
          }
          tot.add(w_prod, w_expo, w_sum, w_s2);
        }
        if (lane == 0) {
          sh.irt.lp[warp] = tot.lp();
          sh.irt.s2[warp] = tot.S2;
        }
        __syncthreads();
      } else {
#ifndef B200_SIMT_EMUL
        double va[IPL], vb[IPL], w_prod, w_sum, w_s2;
        int w_expo;
        // theta gradients of this CTA's block go to its shared memory (index i - base): CTA 0 writes all of g itself, so that
        // everything its advance step reads was written by its own SM (no stale L1 lines)
        irt_block<IPL, 1, true>(sh.irt, cp, y, I, J, rank, a, b, sh.th - (1 + rank * PB), va, vb, w_prod, w_expo, w_sum, w_s2);
        if (lane < IRT_G) {
#pragma unroll
          for (int t = 0; t < IPL; t++) {
            sh.irt.ga[warp][sub + IRT_G * t] = va[t];
            sh.irt.gb[warp][sub + IRT_G * t] = vb[t];
          }
        }
        if (lane == 0) {
          sh.irt.blk[warp][0] = w_prod;
          sh.irt.blk[warp][1] = (double)w_expo;
          sh.irt.blk[warp][2] = w_sum;
          sh.irt.blk[warp][3] = w_s2;
        }
        cluster.sync();
        if (rank == 0) {
          // block partials of all eight CTAs, added in block order (the order of irt_chain_body)
          if (tid < IRT_MAXJ) {
            for (int w = 0; w < IRT_NW; w++) {
              double sa = 0.0, sb = 0.0;
              for (int bk = 0; bk < IRT_NB; bk++) {
                const SoloShared* r = cluster.map_shared_rank(&sh, bk);
                sa += r->irt.ga[w][tid];
                sb += r->irt.gb[w][tid];
              }
              sh.irt.ga[w][tid] = sa;   // (w, tid) of this CTA's own partial was read above, by this thread only
              sh.irt.gb[w][tid] = sb;
            }
          }
          if (tid >= IRT_MAXJ && tid < IRT_MAXJ + IRT_NW) {
            const int w = tid - IRT_MAXJ;
            IrtWarpTotals tot;
            for (int bk = 0; bk < IRT_NB; bk++) {
              const SoloShared* r = cluster.map_shared_rank(&sh, bk);
              tot.add(r->irt.blk[w][0], (int)r->irt.blk[w][1], r->irt.blk[w][2], r->irt.blk[w][3]);
            }
            sh.irt.lp[w] = tot.lp();
            sh.irt.s2[w] = tot.S2;
          }
          for (int i = tid; i < I; i += IRT_NT) {
            const int bk = i / PB;
            const SoloShared* r = cluster.map_shared_rank(&sh, bk);
            v.g[1 + i] = r->th[i - bk * PB];
          }
        }
        __syncthreads();
#endif
      }
      if (rank == 0) irt_finish<true>(sh.irt, cp, I, J, v.g, v.zV);
    }
    __threadfence_block();
    __syncthreads();
    // ---- advance the chain to its next gradient request (CTA 0)
    if (rank == 0) {
      int r = SOLO_GENERIC;
      const int st = sh.chain.state;
      __syncthreads();   // every warp has read the state before warp 0 (advance) may change it
      if (st == ST_LEAF) r = solo_leaf_step(v, cfg, red, &sh.fl);
      if (r == SOLO_GENERIC) {
        if (warp == 0) {
#ifdef B200_SIMT_EMUL
          simt_replica = sh.chain;
          __syncwarp();
          ChainView vr = v;
          vr.s = &simt_replica;
          const int a = advance_out_of_line(vr, &cfg, lane);
          __syncwarp();
          if (lane == 0) sh.chain = simt_replica;
#else
          const int a = advance_out_of_line(v, &cfg, lane);
#endif
          __syncwarp();
          if (lane == 0) sh.more = a == ADV_NEED_GRAD ? 1 : 0;
        }
      } else if (tid == 0) {
        sh.more = 1;
      }
      __threadfence();   // z.q is read by the other CTAs of the cluster after the barrier
    }
    __syncthreads();
    SOLO_CLUSTER_SYNC();
    int more;
#ifndef B200_SIMT_EMUL
    if (CL > 1) more = cluster.map_shared_rank(&sh, 0)->more;
    else
#endif
      more = sh.more;
    if (!more) break;
  }
  SOLO_CLUSTER_SYNC();   // nobody leaves while CTA 0's flag may still be read
  if (rank == 0 && warp == 0) warp_copy_words(gs, &sh.chain, kChainWords, lane);
#undef SOLO_CLUSTER_SYNC
}

size_t irt_solo_smem_bytes(int ld) { return (size_t)6 * ld * sizeof(double); }
bool irt_solo_supported(const ModelData& m) { return (m.N + IRT_NB - 1) / IRT_NB <= IRT_SOLO_MAX_PB; }

// one cluster (cluster_size 8) or one CTA (cluster_size 1) per entry of the compacted list of requesting chains (n_max = the
// host's count of them).  Returns cudaSuccess or the launch error.
cudaError_t launch_irt_solo(const SamplerCfg& cfg, const DeviceChains& d, const ModelData& m, const int* list, const int* n_list, int n_max,
                            int cluster_size, cudaStream_t stream) {
  const int ipl = (m.K + IRT_G - 1) / IRT_G;
  const size_t smem = irt_solo_smem_bytes(cfg.ld);
#ifdef B200_SIMT_EMUL
#define SOLO_LAUNCH(IPL)                                                                                              \
  {                                                                                                                   \
    simt::launch(n_max, IRT_NT, smem, [&] { irt_solo_kernel<IPL, 1>(cfg, d, m.ypk, m.N, m.K, list, n_list); });        \
    return 0;                                                                                                         \
  }
#else
#define SOLO_LAUNCH(IPL)                                                                                              \
  {                                                                                                                   \
    if (cluster_size == IRT_NB) {                                                                                     \
      auto k = irt_solo_kernel<IPL, IRT_NB>;                                                                          \
      cudaError_t e1 = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);               \
      if (e1 != cudaSuccess) return e1;                                                                               \
      cudaLaunchConfig_t lc = {};                                                                                     \
      lc.gridDim = dim3((unsigned)n_max * IRT_NB);                                                                    \
      lc.blockDim = dim3(IRT_NT);                                                                                     \
      lc.dynamicSmemBytes = smem;                                                                                     \
      lc.stream = stream;                                                                                             \
      cudaLaunchAttribute at[1];                                                                                      \
      at[0].id = cudaLaunchAttributeClusterDimension;                                                                 \
      at[0].val.clusterDim.x = IRT_NB;                                                                                \
      at[0].val.clusterDim.y = 1;                                                                                     \
      at[0].val.clusterDim.z = 1;                                                                                     \
      lc.attrs = at;                                                                                                  \
      lc.numAttrs = 1;                                                                                                \
      return cudaLaunchKernelEx(&lc, k, cfg, d, (const unsigned short*)m.ypk, m.N, m.K, list, n_list);                    \
    }                                                                                                                 \
    auto k = irt_solo_kernel<IPL, 1>;                                                                                 \
    cudaError_t e1 = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                 \
    if (e1 != cudaSuccess) return e1;                                                                                 \
    k<<<n_max, IRT_NT, smem, stream>>>(cfg, d, m.ypk, m.N, m.K, list, n_list);                                         \
    return cudaGetLastError();                                                                                        \
  }
#endif
  if (ipl <= 2) SOLO_LAUNCH(2)
  else if (ipl <= 4) SOLO_LAUNCH(4)
  else if (ipl <= 7) SOLO_LAUNCH(7)
  else if (ipl <= 10) SOLO_LAUNCH(10)
  else if (ipl <= 13) SOLO_LAUNCH(13)
  else SOLO_LAUNCH(16)
#undef SOLO_LAUNCH
}

}  // namespace b200
