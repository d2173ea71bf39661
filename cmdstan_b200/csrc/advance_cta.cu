// The lock-step engine's advance step for LONG parameter vectors (D >= kAdvanceCtaD): one CTA of 256 threads per chain instead of
// one warp.  With a warp per chain a vector pass over D = 1204 elements (c4) is five rounds of eight dependent global loads per
// lane and a leaf needs ~29 passes, so the step is bound by memory latency at a quarter of the HBM rate (r02c: 1.3 ms per
// lock-step of 8192 chains beside a 2.9 ms gradient stage).  Here the 256 threads share the element-wise work of the leaf path
// (solo_leaf.cuh: every pass is one round trip) and warp 0 runs the rest of the state machine (advance(), a few times per
// transition).  Same bits as advance_kernel by construction (solo_leaf.cuh explains why), so the engine may pick either.
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
  int need;
};

__global__ void __launch_bounds__(SOLO_NT, 3) advance_cta_kernel(SamplerCfg cfg, DeviceChains d, int C, int* active, int* list, int* n_active) {
  __shared__ __align__(16) AdvanceCtaShared sh;
  B200_DYN_SMEM(red);  // 6 * ld doubles (solo_leaf_step)
  const int c = blockIdx.x;
  if (c >= C) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  ChainView v = make_view(d, cfg, c);
  ChainS* gs = v.s;
  if (warp == 0) warp_copy_words(&sh.chain, gs, kChainWords, lane);
  __syncthreads();
  v.s = &sh.chain;
  const int st = sh.chain.state;
  __syncthreads();  // every warp has read the state before anyone changes it
  int r = SOLO_GENERIC;
  if (st == ST_LEAF) r = solo_leaf_step(v, cfg, red, &sh.fl);
  if (r == SOLO_GENERIC) {
    if (warp == 0) {
      ChainView vw = v;
      B200_SCALARS_BEGIN(vw);   // CPU test rig only: its free-running lanes work on replicas of the scalar block (chains.cuh)
      const int a = advance(vw, cfg, lane);
      B200_SCALARS_END(vw, lane);
      __syncwarp();
      if (lane == 0) sh.need = a == ADV_NEED_GRAD ? 1 : 0;
    }
  } else if (tid == 0) {
    sh.need = 1;
  }
  __syncthreads();
  if (warp == 0) {
    warp_copy_words(gs, &sh.chain, kChainWords, lane);
    if (lane == 0) {
      const int a = sh.need;
      active[c] = a;
      if (a) list[atomicAdd(n_active, 1)] = c;
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
  return 0;
#else
  return cudaGetLastError();
#endif
}

}  // namespace b200
