// The lock-step engine's advance step for LONG parameter vectors (D >= kAdvanceCtaD), in two kernels.
//
// With a warp per chain a vector pass over D = 1204 elements (c4) is five rounds of eight dependent global loads per lane and a
// tree leaf needs ~29 passes, so the step is bound by memory latency (r02f: 1.57 ms per lock-step of 8192 chains beside a 2.6 ms
// gradient stage).  Four out of five requests come from the leaf path of the tree (ST_LEAF), so:
//   1. advance_cta_kernel: one CTA of 256 threads per chain that is in ST_LEAF; its threads share the element-wise work of the
//      leaf (solo_leaf.cuh: every pass is one round trip).  A chain in any other state, or whose subtree just ended, is only
//      MARKED (active[c] = -1): running the rest of the state machine on warp 0 of a 256-thread CTA would idle seven warps out of
//      eight for the longest requests of the step (r02h capture: 68 % of the samples at the final barrier);
//   2. advance_kernel<1> (advance_big.cu) with only_pending = 1: one warp per marked chain, sixteen of them per SM.
// Same bits as the one-kernel step by construction (solo_leaf.cuh explains why), so the engine may pick either.
#define B200_VEC_UNROLL 8
#define B200_WITH_DENSE 0
#include "chains.cuh"
#include "kernels.h"
#include "solo_leaf.cuh"

namespace b200 {

namespace {

struct AdvanceCtaShared {
  ChainS chain;
  SoloFlags fl;
};

__global__ void __launch_bounds__(SOLO_NT, 3) advance_cta_kernel(SamplerCfg cfg, DeviceChains d, int C, int* active, int* list, int* n_active) {
  __shared__ __align__(16) AdvanceCtaShared sh;
  B200_DYN_SMEM(red);  // 6 * ld doubles (solo_leaf_step)
  const int c = blockIdx.x;
  if (c >= C) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  ChainView v = make_view(d, cfg, c);
  ChainS* gs = v.s;
  if (gs->state != ST_LEAF) {  // CTA-uniform: left to the warp-per-chain kernel that follows
    if (tid == 0) active[c] = -1;
    return;
  }
  if (warp == 0) warp_copy_words(&sh.chain, gs, kChainWords, lane);
  __syncthreads();
  v.s = &sh.chain;
  const int r = solo_leaf_step(v, cfg, red, &sh.fl);
  if (warp == 0) {
    warp_copy_words(gs, &sh.chain, kChainWords, lane);
    if (lane == 0) {
      if (r == SOLO_NEED_GRAD) {
        active[c] = 1;
        list[atomicAdd(n_active, 1)] = c;
      } else {
        active[c] = -1;  // the subtree ended: advance() continues from ST_SUBTREE_END in the next kernel
      }
    }
  }
}

}  // namespace

cudaError_t launch_advance_cta(const SamplerCfg& cfg, const DeviceChains& d, int C, int* active, int* list, int* n_active, cudaStream_t stream) {
  const size_t smem = (size_t)6 * cfg.ld * sizeof(double);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e1 = cudaFuncSetAttribute(advance_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e1 != 0) return e1;
    configured = smem;
  }
  B200_LAUNCH(advance_cta_kernel, C, SOLO_NT, smem, stream, cfg, d, C, active, list, n_active);
#ifdef B200_SIMT_EMUL
  simt::launch((C + kChainsPerBlock - 1) / kChainsPerBlock, 32 * kChainsPerBlock, 0, [&] { advance_kernel<1>(cfg, d, C, active, list, n_active, 1); });
  return 0;
#else
  launch_advance_big(cfg, d, C, active, list, n_active, stream, 1);
  return cudaGetLastError();
#endif
}

}  // namespace b200
