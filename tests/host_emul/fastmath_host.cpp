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

// ---- round 2: table-driven exp and the product form of the logistic log-likelihood
extern "C" void fastmath_exp_t(const double* x, int n, double* out) {
  for (int i = 0; i < n; i++) out[i] = b200::fast_exp_t(x[i], b200::kExp2Tab32);
}
extern "C" void fastmath_rcp3(const double* x, int n, double* out) {
  for (int i = 0; i < n; i++) out[i] = b200::fast_rcp3(x[i]);
}
// every observation through its own accumulator: the per-term value and derivative
extern "C" void fastmath_logit_terms(const double* eta, const double* yd, int n, double* lp, double* d) {
  for (int i = 0; i < n; i++) {
    b200::LogitAcc a;
    b200::logit_acc_init(a);
    d[i] = b200::logit_term(a, eta[i], yd[i], b200::kExp2Tab32);
    b200::logit_acc_fold(a);
    lp[i] = b200::logit_acc_value(a);
  }
}
// all observations through ONE accumulator, folded every `fold_every` terms; `n_pad` padding observations (eta == 0) are mixed in
// and taken out again by logit_acc_value
extern "C" double fastmath_logit_sum(const double* eta, const double* yd, int n, int fold_every, int n_pad, double* d) {
  b200::LogitAcc a;
  b200::logit_acc_init(a);
  int pads = 0;
  for (int i = 0; i < n; i++) {
    d[i] = b200::logit_term(a, eta[i], yd[i], b200::kExp2Tab32);
    if (pads < n_pad && i % 3 == 0) {
      b200::logit_term(a, (i & 1) ? 0.0 : -0.0, (i & 2) ? 1.0 : -1.0, b200::kExp2Tab32);
      pads++;
    }
    if ((i + 1) % fold_every == 0) b200::logit_acc_fold(a);
  }
  for (; pads < n_pad; pads++) b200::logit_term(a, 0.0, 1.0, b200::kExp2Tab32);
  b200::logit_acc_fold(a);
  return b200::logit_acc_value(a, n_pad);
}
