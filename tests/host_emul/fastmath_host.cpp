// TEST-ONLY: compiles the branch-free link functions of the data-heavy gradient kernels (cmdstan_b200/csrc/fastmath.cuh) for
// the host so that their accuracy claim (relative error <= 4e-16 per term) and their edge cases (the +-20 cut-offs of
// bernoulli_logit_glm_lpmf, clamps, NaN / inf propagation) can be checked against long-double references without a GPU.
// The device intrinsics are replaced by bit-exact host equivalents; rcp.approx.ftz.f64 (a ~20-bit reciprocal seed) by a
// float-precision seed, which the two Newton steps that follow refine to the same last-bit accuracy.
#include <cmath>
#include <cstdint>
#include <cstring>

#define __device__
#define __forceinline__ inline
static inline int __double2loint(double x) { uint64_t b; std::memcpy(&b, &x, 8); return (int)(uint32_t)(b & 0xffffffffu); }
static inline int __double2hiint(double x) { uint64_t b; std::memcpy(&b, &x, 8); return (int)(uint32_t)(b >> 32); }
static inline double __hiloint2double(int hi, int lo) { uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double x; std::memcpy(&x, &b, 8); return x; }
#define B200_FASTMATH_HOST 1
#include "../../cmdstan_b200/csrc/fastmath.cuh"

extern "C" void fastmath_link(int link, const double* eta, const double* yd, int n, double* lp, double* d) {
  for (int i = 0; i < n; i++) {
    if (link == 0) b200::link_eval<0>(eta[i], yd[i], lp[i], d[i]);
    else b200::link_eval<1>(eta[i], yd[i], lp[i], d[i]);
  }
}
extern "C" void fastmath_exp(const double* x, int n, double* out) {
  for (int i = 0; i < n; i++) out[i] = b200::fast_exp(x[i]);
}
