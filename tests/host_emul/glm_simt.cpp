// TEST-ONLY: the FP64 tensor-core GLM gradient kernels (cmdstan_b200/csrc/glm_dmma.cu: v2 block-synchronous, v3
// warp-asynchronous, v4 three stages + shuffle exchange) executed on host threads by the SIMT stand-in under
// tests/host_emul/simt, through the product's own launcher (variant / template dispatch included).  Used by
// tests/test_glm_kernels_simt.py: per row-group partials against a numpy restatement of the GLM likelihood, v2 / v3 (validated
// on hardware) as controls for the rig, v4 (not yet run on hardware) as the subject; built a second time with
// -fsanitize=thread to look for shared-memory races in the refill protocols.
#define B200_SIMT_EMUL 1
#define B200_FASTMATH_HOST 1
#include <cuda_runtime.h>  // tests/host_emul/simt/cuda_runtime.h

#include "../../cmdstan_b200/csrc/glm_dmma.cu"

// X: [N][ldx] (ldx = Dk), y: [N], Q: [C][ldq], list: n_list chain indices or null.  Outputs: Gp [P][cap][Dk], LPp [P][cap].
extern "C" int glm_simt_run(int variant, int logistic, int N, int D, int Dk, int ldq, int C, int n_list, int P, const double* X, const int* y,
                            const double* Q, const int* list, double* Gp, double* LPp) {
  b200::ModelData m{};
  m.kind = logistic ? b200::MODEL_LOGISTIC : b200::MODEL_POISSON;
  m.D = D;
  m.N = N;
  m.K = D - 1;
  m.ldx = Dk;
  m.X = X;
  m.yi = y;
  b200::GradBatch b{};
  b.Q = Q;
  b.C = C;
  b.ld = ldq;
  b.list = list;
  int nl = n_list;
  b.n_list = list ? &nl : nullptr;
  if (!list) b.C = n_list;
  b200::GlmWorkspace ws;
  ws.Gp = Gp;
  ws.LPp = LPp;
  ws.P = P;
  ws.C = C;
  ws.Dp = Dk;
  ws.variant = variant;
  return b200::launch_glm_dmma(m, b, ws, nullptr);
}
extern "C" size_t glm_simt_smem_bytes(int variant, int Dk) {
  if (variant == 3) return b200::glm_tc4_smem_bytes(Dk, b200::glm_tc4_smem_bytes(Dk, 3) <= 232448 ? 3 : 2);
  if (variant == 2) return b200::glm_tc3_smem_bytes(Dk);
  if (variant == 4) return b200::glm_tc5_smem_bytes(Dk);
  if (variant == 5) return b200::glm_tc6_smem_bytes(Dk);
  if (variant == 6) return b200::glm_tc7_smem_bytes(Dk);
  return b200::glm_tc_smem_bytes(Dk);
}

#ifdef GLM_SIMT_MAIN
extern "C" int glm_simt_grad(int variant, int logistic, int N, int D, int ldq, int C, int sm_count, const double* X, const int* y, const double* Q,
                             const int* active, const int* list, int n_list, double* G, double* zV);  // glm_simt_stage.cpp
// stand-alone form for ThreadSanitizer (g++ -fsanitize=thread -DGLM_SIMT_MAIN): synthetic data, one launch, a checksum
#include <random>
int main(int argc, char** argv) {
  if (argc < 8) {
    std::fprintf(stderr, "usage: %s variant logistic N D C n_list P\n", argv[0]);
    return 2;
  }
  const int variant = std::atoi(argv[1]), logistic = std::atoi(argv[2]), N = std::atoi(argv[3]), D = std::atoi(argv[4]), C = std::atoi(argv[5]),
            n_list = std::atoi(argv[6]), P = std::atoi(argv[7]);
  const int Dk = (D + 7) / 8 * 8, ldq = (D + 1) / 2 * 2;
  std::mt19937_64 gen(7);
  std::normal_distribution<double> nd;
  std::vector<double> X((size_t)N * Dk, 0.0), Q((size_t)C * ldq, 0.0), Gp((size_t)P * C * Dk, 0.0), LPp((size_t)P * C, 0.0);
  std::vector<int> y(N), list(n_list);
  for (int i = 0; i < N; i++) {
    X[(size_t)i * Dk] = 1.0;
    for (int k = 1; k < D; k++) X[(size_t)i * Dk + k] = nd(gen);
    y[i] = (int)(gen() % (logistic ? 2 : 5));
  }
  for (int c = 0; c < C; c++)
    for (int k = 0; k < D; k++) Q[(size_t)c * ldq + k] = 0.3 * nd(gen);
  for (int s = 0; s < n_list; s++) list[s] = (int)((s * 7919LL) % C);
  if (variant == 0) {  // the register-tiled DFMA kernel + reduction + finishing kernel of glm.cu (every chain requests)
    std::vector<double> G((size_t)C * ldq, 0.0), zV(C, 0.0);
    glm_simt_grad(0, logistic, N, D, ldq, C, P, X.data(), y.data(), Q.data(), nullptr, nullptr, 0, G.data(), zV.data());
    double sum = 0;
    for (int c = 0; c < C; c++) sum -= zV[c];
    std::printf("lp checksum %.17g\n", sum);
    return 0;
  }
  glm_simt_run(variant, logistic, N, D, Dk, ldq, C, n_list, P, X.data(), y.data(), Q.data(), list.data(), Gp.data(), LPp.data());
  double sum = 0;
  for (int p = 0; p < P; p++)
    for (int s = 0; s < n_list; s++) sum += LPp[(size_t)p * C + s];
  std::printf("lp checksum %.17g\n", sum);
  return 0;
}
#endif
