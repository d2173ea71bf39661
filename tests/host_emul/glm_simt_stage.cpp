// TEST-ONLY, second translation unit of the GLM kernel rig (see glm_simt.cpp): cmdstan_b200/csrc/glm.cu -- the register-tiled
// DFMA partial kernel, the reduction over row groups and the finishing kernel (priors, signs) -- on host threads.
#define B200_SIMT_EMUL 1
#define B200_FASTMATH_HOST 1
#include <cuda_runtime.h>  // tests/host_emul/simt/cuda_runtime.h

#include "../../cmdstan_b200/csrc/glm.cu"

// the whole gradient stage of the lock-step engine for the GLMs (launch_glm_grad: partial kernel of the chosen variant -- 0 is the
// register-tiled DFMA kernel of glm.cu --, reduction over row groups, priors and signs): zV[c] = -lp, G[c] = -grad lp
extern "C" int glm_simt_grad(int variant, int logistic, int N, int D, int ldq, int C, int sm_count, const double* X, const int* y, const double* Q,
                             const int* active, const int* list, int n_list, double* G, double* zV) {
  b200::ModelData m{};
  m.kind = logistic ? b200::MODEL_LOGISTIC : b200::MODEL_POISSON;
  m.D = D; m.N = N; m.K = D - 1; m.ldx = (D + 7) & ~7; m.X = X; m.yi = y;
  b200::GlmWorkspace ws;
  const size_t bytes = b200::glm_workspace_bytes(N, D, C, sm_count, &ws);
  std::vector<double> buf(bytes / sizeof(double), std::nan(""));
  b200::glm_workspace_bind(&ws, buf.data());
  ws.variant = variant;
  b200::GradBatch b{};
  b.Q = Q; b.G = G; b.zV = zV; b.active = active; b.C = C; b.ld = ldq;
  int nl = n_list;
  if (list) { b.list = list; b.n_list = &nl; }
  return b200::launch_glm_grad(m, b, ws, nullptr, nullptr, true, nullptr);
}
