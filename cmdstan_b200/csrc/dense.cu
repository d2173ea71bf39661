// The chain kernels with the dense-metric code paths AND the diagnostics bookkeeping compiled in (nuts.cuh B200_WITH_DENSE = 1):
// hmc_nuts_dense_e[_adapt] and any run with store_diagnostics.
// Same source as engine.cu's kernels (chains.cuh / nuts.cuh); kept in their own translation unit so that the diag_e / unit_e
// kernels -- the path the benchmarks run -- do not carry the dense branches.  Built with --fmad=false like engine.cu.
#define B200_WITH_DENSE 1
#include "chains.cuh"

namespace b200 {

void launch_advance_dense(const SamplerCfg& cfg, const DeviceChains& d, int C, int* active, int* list, int* n_active, cudaStream_t stream) {
  const int blocks = (C + kChainsPerBlock - 1) / kChainsPerBlock;
  advance_kernel<2><<<blocks, 32 * kChainsPerBlock, 0, stream>>>(cfg, d, C, active, list, n_active);
}

cudaError_t launch_fused_dense(const SamplerCfg& cfg, const ModelData& md, const DeviceChains& d, int C, bool smem_state, int warps_per_block,
                               int chain_words, size_t smem_bytes, cudaStream_t stream) {
  const int grid = (C + warps_per_block - 1) / warps_per_block;
  if (smem_state) {
    auto k = fused_chain_kernel<true, 2>;
    cudaError_t r = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (r != cudaSuccess) return r;
    k<<<grid, 32 * warps_per_block, smem_bytes, stream>>>(cfg, md, d, C, warps_per_block, chain_words);
  } else {
    fused_chain_kernel<false, 2><<<grid, 32 * warps_per_block, 0, stream>>>(cfg, md, d, C, warps_per_block, 0);
  }
  return cudaGetLastError();
}

}  // namespace b200
