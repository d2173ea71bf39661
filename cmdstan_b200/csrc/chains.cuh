// Device-side chain storage of the engine and the lock-step engine's advance kernel (shared by engine.cu and
// advance_big.cu, which builds the same kernel with the vector loops unrolled for long parameter vectors).
#pragma once
#include <cuda_runtime.h>

#include "nuts.cuh"

namespace b200 {

constexpr int kChainsPerBlock = 4;  // 4 warps per CTA
constexpr int kCounterRing = 16;    // lock-step engine: host checks for completion every kCounterRing steps
constexpr int kAdvanceBigD = 64;    // parameter vectors at least this long use the unrolled advance kernel (advance_big.cu)

struct DeviceChains {
  ChainS* cs;
  double *q, *g, *zV, *slab, *draws;
  const double* user_q0;
  size_t slab_stride, draws_stride;
  double* mat;        // dense metric: [C][3][D][D] (inverse metric, its Cholesky factor, Welford m2), else nullptr
  size_t mat_stride;  // 3 D D
};

__device__ __forceinline__ ChainView make_view(const DeviceChains& d, const SamplerCfg& cfg, int c) {
  ChainView v;
  v.s = d.cs + c;
  v.q = d.q + (size_t)c * cfg.ld;
  v.g = d.g + (size_t)c * cfg.ld;
  v.zV = d.zV + c;
  v.v = d.slab + (size_t)c * d.slab_stride;
  v.draws = d.draws ? d.draws + (size_t)c * d.draws_stride : nullptr;
  v.user_q0 = d.user_q0 ? d.user_q0 + (size_t)c * cfg.ld : nullptr;
  v.mat = d.mat ? d.mat + (size_t)c * d.mat_stride : nullptr;
  return v;
}

// ---- copy helpers: a chain's scalar block / vectors between global and shared memory (whole warp, 8-byte words)
__device__ __forceinline__ void warp_copy_words(void* dst, const void* src, int n_words, int lane) {
  unsigned long long* d = reinterpret_cast<unsigned long long*>(dst);
  const unsigned long long* s = reinterpret_cast<const unsigned long long*>(src);
#pragma unroll 1
  for (int i = lane; i < n_words; i += 32) d[i] = s[i];
}
static_assert(sizeof(ChainS) % 8 == 0, "ChainS is copied as 8-byte words");
constexpr int kChainWords = (int)(sizeof(ChainS) / 8);

#ifdef B200_SIMT_EMUL  // CPU test-suite only (tests/host_emul/simt): the chain kernels run on host threads
#define B200_DYN_SMEM_WORDS(name) unsigned long long* name = reinterpret_cast<unsigned long long*>(simt::dyn_smem())
// The device's lanes read and write the chain's warp-uniform scalar block (ChainS) in lockstep; the rig's lanes are free-running
// threads, so there each lane works on a private replica (all replicas evolve identically; the random stream is only ever
// touched by lane 0, whose replica is the one stored back)
#define B200_SCALARS_BEGIN(v)       \
  ChainS* const simt_home = (v).s;  \
  ChainS simt_replica = *simt_home; \
  __syncwarp();                     \
  (v).s = &simt_replica
#define B200_SCALARS_END(v, lane)                \
  __syncwarp();                                  \
  if ((lane) == 0) *simt_home = simt_replica;    \
  (v).s = simt_home;                             \
  __syncwarp()
#else
#define B200_DYN_SMEM_WORDS(name) extern __shared__ __align__(16) unsigned long long name[]
#define B200_SCALARS_BEGIN(v)
#define B200_SCALARS_END(v, lane)
#endif

// ---- lock-step engine: advance every chain to its next gradient request
// and collect the requesting chains: active[c] flag + compacted index list (order irrelevant: a chain's gradient
// depends on its own q only).  The chain's scalar block is staged in shared memory for the launch (one coalesced
// load + store instead of dozens of dependent global accesses).
// only_pending != 0: second half of a two-kernel advance step (advance_cta.cu): only the chains the first kernel marked with
// active[c] == -1 are advanced, the others have already been served.
template <int TAG>
__global__ void __launch_bounds__(32 * kChainsPerBlock) advance_kernel(SamplerCfg cfg, DeviceChains d, int C, int* active,
                                                                          int* list, int* n_active, int only_pending = 0) {
  __shared__ __align__(16) ChainS sh[kChainsPerBlock];
  const int w = threadIdx.x >> 5;
  const int c = blockIdx.x * kChainsPerBlock + w;
  if (c >= C) return;
  if (only_pending && active[c] != -1) return;
  const int lane = threadIdx.x & 31;
  ChainView v = make_view(d, cfg, c);
  ChainS* gs = v.s;
  warp_copy_words(&sh[w], gs, kChainWords, lane);
  __syncwarp();
  v.s = &sh[w];
  B200_SCALARS_BEGIN(v);
  const int r = advance(v, cfg, lane);
  B200_SCALARS_END(v, lane);
  __syncwarp();
  warp_copy_words(gs, &sh[w], kChainWords, lane);
  if (lane == 0) {
    const int a = r == ADV_NEED_GRAD ? 1 : 0;
    active[c] = a;
    if (a) list[atomicAdd(n_active, 1)] = c;
  }
}


// ---- fused engine: advance + inlined model gradient, looped until the chain is idle.
// SMEM = true: the chain's whole state (scalars, z.q, z.g, vector slab) lives in shared memory for the duration of the
// launch (small models: 7 KB per chain for eight schools) and is written back at the end, so the state machine's
// many dependent accesses cost shared-memory latency instead of L2/HBM latency.
constexpr int kFusedMaxWarps = 7;  // 4 blocks x 7 warps = 28 resident chains per SM (4096 chains = one wave on 148 SMs)
template <bool SMEM, int TAG>
__global__ void __launch_bounds__(32 * kFusedMaxWarps, 4) fused_chain_kernel(SamplerCfg cfg, ModelData md, DeviceChains d, int C, int warps_per_block, int chain_smem_words) {
  B200_DYN_SMEM_WORDS(fused_smem);
  const int w = threadIdx.x >> 5;
  const int c = blockIdx.x * warps_per_block + w;
  if (c >= C) return;
  const int lane = threadIdx.x & 31;
  ChainView v = make_view(d, cfg, c);
  ChainView g = v;  // where the state lives between launches
  const int slab_words = (int)d.slab_stride;
  if (SMEM) {
    unsigned long long* base = fused_smem + (size_t)w * chain_smem_words;
    v.s = reinterpret_cast<ChainS*>(base);
    double* vec = reinterpret_cast<double*>(base + kChainWords);
    v.q = vec;
    v.g = vec + cfg.ld;
    v.zV = vec + 2 * cfg.ld;
    v.v = vec + 2 * cfg.ld + 2;
    warp_copy_words(v.s, g.s, kChainWords, lane);
    warp_copy_words(v.q, g.q, cfg.ld, lane);
    warp_copy_words(v.g, g.g, cfg.ld, lane);
    warp_copy_words(v.v, g.v, slab_words, lane);
    if (lane == 0) *v.zV = *g.zV;
    __syncwarp();
  }
  B200_SCALARS_BEGIN(v);
  while (advance(v, cfg, lane) == ADV_NEED_GRAD) {
    __syncwarp();
    const double lp = lp_grad_small(md, v.q, v.g, lane);
    for (int i = lane; i < cfg.D; i += 32) v.g[i] = -v.g[i];
    *v.zV = -lp;
    __syncwarp();
  }
  B200_SCALARS_END(v, lane);
  if (SMEM) {
    __syncwarp();
    warp_copy_words(g.s, v.s, kChainWords, lane);
    warp_copy_words(g.q, v.q, cfg.ld, lane);
    warp_copy_words(g.g, v.g, cfg.ld, lane);
    warp_copy_words(g.v, v.v, slab_words, lane);
    if (lane == 0) *g.zV = *v.zV;
  }
}

// advance_cta.cu: one CTA per chain for very long parameter vectors (bit-identical results)
constexpr int kAdvanceCtaD = 512;
constexpr int kAdvanceCtaMaxLd = 4608;  // 6 * ld doubles of shared memory must fit (216 KB)
cudaError_t launch_advance_cta(const SamplerCfg& cfg, const DeviceChains& d, int C, int* active, int* list, int* n_active, cudaStream_t stream);
// advance_big.cu: the same kernel built with B200_VEC_UNROLL = 8 (bit-identical results)
void launch_advance_big(const SamplerCfg& cfg, const DeviceChains& d, int C, int* active, int* list, int* n_active, cudaStream_t stream,
                        int only_pending = 0);
// irt.cu: latency form for the stragglers of an IRT run (one cluster of 8 CTAs, or one CTA, per remaining chain)
bool irt_solo_supported(const ModelData& m);
cudaError_t launch_irt_solo(const SamplerCfg& cfg, const DeviceChains& d, const ModelData& m, const int* list, const int* n_list, int n_max,
                            int cluster_size, cudaStream_t stream);
// dense.cu: the kernels built with B200_WITH_DENSE = 1 (metric = dense_e)
void launch_advance_dense(const SamplerCfg& cfg, const DeviceChains& d, int C, int* active, int* list, int* n_active, cudaStream_t stream);
cudaError_t launch_fused_dense(const SamplerCfg& cfg, const ModelData& md, const DeviceChains& d, int C, bool smem_state, int warps_per_block,
                               int chain_words, size_t smem_bytes, cudaStream_t stream);

}  // namespace b200
