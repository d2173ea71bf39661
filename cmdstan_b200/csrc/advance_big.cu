// The lock-step engine's advance kernel for long parameter vectors: same source as engine.cu's (chains.cuh / nuts.cuh), with
// the per-lane vector loops unrolled 8x so that a warp keeps 8 elements per lane in flight (see nuts.cuh B200_VEC_UNROLL).
// Built with --fmad=false like engine.cu: the two kernels must round identically.
#define B200_VEC_UNROLL 8
#define B200_WITH_DENSE 0
#include "chains.cuh"

namespace b200 {

void launch_advance_big(const SamplerCfg& cfg, const DeviceChains& d, int C, int* active, int* list, int* n_active, cudaStream_t stream,
                        int only_pending) {
  const int blocks = (C + kChainsPerBlock - 1) / kChainsPerBlock;
  advance_kernel<1><<<blocks, 32 * kChainsPerBlock, 0, stream>>>(cfg, d, C, active, list, n_active, only_pending);
}

}  // namespace b200
