// Host-callable launchers of the cross-chain gradient kernels (lock-step engine).  Every launcher computes, for each
// chain c with active[c] != 0 (active == nullptr: all), the potential and its gradient at Q[c][:] :
//     zV[c] = -lp(Q[c]) ,  G[c][:] = -grad lp(Q[c])          (base_hamiltonian::update_potential_gradient,
//                                                              S/mcmc/hmc/hamiltonians/base_hamiltonian.hpp:61-70)
// with lp = log_prob_propto_jacobian.  Results for a chain depend only on that chain's q (fixed reduction order),
// so a chain's draws do not depend on which other chains share the batch
// (the invariant of ST/unit/services/sample/hmc_nuts_diag_e_adapt_parallel_match_test.cpp:73-153).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "models.cuh"

// Kernel launches and dynamic shared memory of the gradient kernels go through these two macros so that the CPU test-suite can
// run the same kernels on host threads (tests/host_emul/simt); in the product they are the plain CUDA constructs.
#ifdef B200_SIMT_EMUL
#define B200_DYN_SMEM(name) double* name = reinterpret_cast<double*>(simt::dyn_smem())
#define B200_LAUNCH(kernel, grid, block, smem_bytes, stream, ...) simt::launch(grid, block, smem_bytes, [&] { kernel(__VA_ARGS__); })
#else
#define B200_DYN_SMEM(name) extern __shared__ __align__(128) double name[]
#define B200_LAUNCH(kernel, grid, block, smem_bytes, stream, ...) kernel<<<grid, block, smem_bytes, stream>>>(__VA_ARGS__)
#endif

namespace b200 {

typedef int (*allreduce_fn)(void* device_buf, int64_t count, void* cuda_stream, void* user);

struct GradBatch {
  const double* Q;     // [C][ld]
  double* G;           // [C][ld]
  double* zV;          // [C]
  const int* active;   // [C] or nullptr
  int C, ld;
  const int* list = nullptr;    // compacted indices of the chains with active[c] != 0 (any order), or nullptr
  const int* n_list = nullptr;  // device count of entries in `list`
  int hint_active = -1;         // host's last known number of requesting chains (-1: unknown); only a scheduling hint
};

// ---- GLMs (logistic / poisson): glm.cu -------------------------------------------------------------------------
struct GlmWorkspace {
  double* Gp = nullptr;   // [P][C][Dp]   per row-group partial gradients (second index: slot in the chain list)
  double* LPp = nullptr;  // [P][C]       per row-group partial log-likelihood
  double* R = nullptr;    // [C][1 + Dp]  reduced (lp, gradient): the all-reduce payload for row sharding
  int P = 0, C = 0, Dp = 0;
  int variant = 0;        // 0: register-tiled DFMA kernel, 1: DMMA kernel v2 (block-synchronous), 2..6: other DMMA formulations, 7: INT8 (glm_i8.cu)
  void* i8 = nullptr;     // variant 7: state of the Ozaki-split INT8 kernel (digit-plane image of X, column scales), or nullptr
};
// glm_i8.cu: logistic regression on tcgen05 kind::i8.  glm_i8_create returns nullptr when the formulation does not apply.
void* glm_i8_create(const ModelData& m, int Dk, int cap, int sm_count, int max_row_groups, cudaStream_t stream);
void glm_i8_destroy(void* state);
int glm_i8_row_groups(const void* state);
int launch_glm_i8(const ModelData& m, const GradBatch& b, const GlmWorkspace& ws, cudaStream_t stream);
size_t glm_workspace_bytes(int N, int D, int C, int sm_count, GlmWorkspace* ws);  // fills P, C, Dp (= D padded to 8); returns total bytes
void glm_workspace_bind(GlmWorkspace* ws, void* base);
int glm_max_dim();
// returns the number of kernels launched, or -(code) when the all-reduce callback of a row-sharded run reported failure `code`
int launch_glm_grad(const ModelData& m, const GradBatch& b, const GlmWorkspace& ws, allreduce_fn ar, void* ar_user,
                    bool add_prior, cudaStream_t stream);

// ---- IRT 2PL: irt.cu -------------------------------------------------------------------------------------------
int irt_max_items();
// host: responses y [I][J] (0 / 1) -> ModelData::ypk layout [I][8] (item j = s + 8 t is bit t of word s: the 16 bits a lane of the
// gradient kernels needs for one person, so that a response costs no load and no branch in the inner loop)
inline void irt_pack_responses(const signed char* y, int I, int J, unsigned short* out) {
  for (int i = 0; i < I; i++)
    for (int s = 0; s < 8; s++) {
      unsigned v = 0;
      for (int t = 0; s + 8 * t < J; t++) v |= (unsigned)(y[(size_t)i * J + s + 8 * t] != 0) << t;
      out[(size_t)i * 8 + s] = (unsigned short)v;
    }
}
int launch_irt_grad(const ModelData& m, const GradBatch& b, cudaStream_t stream);

}  // namespace b200
