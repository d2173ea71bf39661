// EXPERIMENTAL (not validated on hardware yet; reached only with B200NUTS_IRT_TAIL set): the straggler path of the lock-step
// engine for the IRT model (profiles/README.md r01e: the number of lock-steps of a run is the leapfrog count of its slowest
// chain, and once only a handful of chains still request gradients every step costs two kernel launches plus single-chain
// latency, ~0.2 ms, for a few microseconds of useful work per SM).
//
// irt_tail_kernel: one CTA per still-requesting chain runs that chain TO THE END of the current b200nuts_run call without
// returning to the host: gradient (the whole CTA, irt_chain_body) -> advance (warp 0, the same state machine as
// advance_kernel) -> gradient -> ... until the chain is idle.  No launches, no host synchronisation, no other chain to wait for.
//
// Bit-identity: a chain may be served by the cross-chain gradient kernel for some leapfrogs and by this kernel for the rest,
// and which one depends on the other chains of the batch, so the two must round identically.  Both kernels of this translation
// unit are built from the same irt_chain_body with --fmad=false (no contraction choices left to the compiler), and the engine
// uses irt_grad_kernel_nf -- not irt.cu's kernel -- for every gradient of a run in which the tail path is enabled.
#define B200_VEC_UNROLL 8
#define B200_WITH_DENSE 0
#include "chains.cuh"
#include "irt_body.cuh"

namespace b200 {

template <int IPL>
__global__ void __launch_bounds__(IRT_NT, 1) irt_grad_kernel_nf(const signed char* __restrict__ y, int I, int J, int D,
                                                                const double* __restrict__ Q, double* __restrict__ G,
                                                                double* __restrict__ zV, const int* __restrict__ active,
                                                                const int* __restrict__ list, const int* __restrict__ n_list, int ld) {
  int c = blockIdx.x;
  if (list) {
    if (c >= *n_list) return;
    c = list[c];
  } else if (active && !active[c]) {
    return;
  }
  irt_chain_body<IPL, 1>(y, I, J, D, Q + (size_t)c * ld, G + (size_t)c * ld, zV + c);
}

template <int IPL>
__global__ void __launch_bounds__(IRT_NT, 1) irt_tail_kernel(SamplerCfg cfg, DeviceChains d, const signed char* __restrict__ y, int I, int J,
                                                             const int* __restrict__ list, const int* __restrict__ n_list) {
  __shared__ __align__(16) ChainS sh;
  __shared__ int s_more;
  if ((int)blockIdx.x >= *n_list) return;
  const int c = list[blockIdx.x];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  ChainView v = make_view(d, cfg, c);
  ChainS* gs = v.s;
  if (warp == 0) warp_copy_words(&sh, gs, kChainWords, lane);
  __syncthreads();
  v.s = &sh;
#ifdef B200_SIMT_EMUL  // the rig's lanes are free-running: warp 0's lanes each work on a replica of the scalar block (chains.cuh)
  ChainS simt_replica = sh;
  __syncthreads();
  if (warp == 0) v.s = &simt_replica;
#endif
  for (;;) {
    // the chain is waiting for the gradient at z.q (written by advance, possibly earlier in this very kernel: coherent loads)
    irt_chain_body<IPL, 1, true>(y, I, J, cfg.D, v.q, v.g, v.zV);
    __threadfence_block();
    __syncthreads();
    if (warp == 0) {
      const int r = advance(v, cfg, lane);
      __syncwarp();
      if (lane == 0) s_more = r == ADV_NEED_GRAD ? 1 : 0;
    }
    __threadfence_block();
    __syncthreads();
    if (!s_more) break;
  }
#ifdef B200_SIMT_EMUL
  if (threadIdx.x == 0) sh = simt_replica;
  __syncthreads();
#endif
  if (warp == 0) warp_copy_words(gs, &sh, kChainWords, lane);
}

#define IRT_NF_DISPATCH(ipl, STMT)   \
  if ((ipl) <= 2) { constexpr int IPL = 2; STMT; }        \
  else if ((ipl) <= 4) { constexpr int IPL = 4; STMT; }   \
  else if ((ipl) <= 7) { constexpr int IPL = 7; STMT; }   \
  else if ((ipl) <= 10) { constexpr int IPL = 10; STMT; } \
  else if ((ipl) <= 13) { constexpr int IPL = 13; STMT; } \
  else { constexpr int IPL = 16; STMT; }

int launch_irt_grad_nf(const ModelData& m, const GradBatch& b, cudaStream_t stream) {
  const int ipl = (m.K + IRT_G - 1) / IRT_G;
  IRT_NF_DISPATCH(ipl, IRT_KERNEL_LAUNCH(irt_grad_kernel_nf<IPL>, b.C, stream, m.y8, m.N, m.K, m.D, b.Q, b.G, b.zV, b.active, b.list, b.n_list, b.ld))
  return 1;
}

// one CTA per entry of the compacted list of requesting chains (n_max = the host's count of them)
void launch_irt_tail(const SamplerCfg& cfg, const DeviceChains& d, const ModelData& m, const int* list, const int* n_list, int n_max,
                     cudaStream_t stream) {
  const int ipl = (m.K + IRT_G - 1) / IRT_G;
  IRT_NF_DISPATCH(ipl, IRT_KERNEL_LAUNCH(irt_tail_kernel<IPL>, n_max, stream, cfg, d, m.y8, m.N, m.K, list, n_list))
}

}  // namespace b200
