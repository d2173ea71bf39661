// Branch-free double-precision building blocks for the per-observation link functions of the data-heavy gradient
// kernels (glm_dmma.cu, irt.cu).  On sm_100 scalar FP64 shares one pipe with the FP64 tensor path (DMMA), and a
// latency-bound library exp/log1p/division sequence (~75 FP64 instructions, several divergent branches) was 45 % of the
// GLM kernel's time (profiles/README.md r01a); these versions are straight-line (full ILP across the elements a lane
// owns) and ~45 FP64 instructions per logistic term.  Relative error: fast_exp <= 4e-16, a logistic term <= 1e-15 (measured on the
// host build, tests/test_fastmath_host.py; gate: 1e-10 vs the reference on sums of such terms).
#pragma once
#ifndef B200_FASTMATH_HOST   // tests/host_emul/fastmath_host.cpp builds this header for the host (accuracy tests without a GPU)
#include <cuda_runtime.h>
#endif

namespace b200 {

// ---- branch-free double-precision building blocks -----------------------------------------------------------------
// exp(x) for |x| <= 700 (callers clamp): x = k ln2 + r, |r| <= ln2/2, Taylor degree 12 (remainder 1.7e-16), 2^k applied
// to the exponent field with integer arithmetic.
__device__ __forceinline__ double fast_exp(double x) {
  const double kMagic = 6755399441055744.0;  // 1.5 * 2^52: rounds to nearest integer in the low word
  const double t = fma(x, 1.4426950408889634074, kMagic);
  const double kf = t - kMagic;
  const int k = __double2loint(t);
  double r = fma(kf, -6.93147180369123816490e-01, x);   // ln2 high part (fdlibm split)
  r = fma(kf, -1.90821492927058770002e-10, r);          // ln2 low part
  double p = 1.0 / 479001600.0;                         // 1/12!
  p = fma(p, r, 1.0 / 39916800.0);
  p = fma(p, r, 1.0 / 3628800.0);
  p = fma(p, r, 1.0 / 362880.0);
  p = fma(p, r, 1.0 / 40320.0);
  p = fma(p, r, 1.0 / 5040.0);
  p = fma(p, r, 1.0 / 720.0);
  p = fma(p, r, 1.0 / 120.0);
  p = fma(p, r, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// 1/x for x in [1, 16]: hardware seed (20 bits) + two Newton steps
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
#ifdef B200_FASTMATH_HOST
  r = (double)(1.0f / (float)x);  // stands in for the ~20-bit hardware seed
#else
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#endif
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}

// per-observation log-likelihood term and d(term)/d(eta).  `yd` is 2y-1 (logistic) or y (poisson) as a double.
//   logistic: bernoulli_logit_glm_lpmf.hpp:121-126 (lp), :137-142 (derivative), cutoffs +-20
//   poisson : poisson_log_glm_lpmf.hpp:117-118, :127-132
template <int LINK>
__device__ __forceinline__ void link_eval(double eta, double yd, double& lpi, double& d) {
  if (LINK == 0) {
    const double t = yd * eta;
    const double a = fabs(t);
    const double em = fast_exp(-fmin(a, 700.0));  // exp(-|t|) in (0, 1]
    // log1p(em) = kln2 + 2 atanh(z), z = (f-1)/(f+1), f = (1+em) / (big ? 2 : 1)  =>  |z| <= 0.1716
    const bool big = em > 0.41421356237309503;
    const double num = em - (big ? 1.0 : 0.0);
    const double den = em + (big ? 3.0 : 2.0);
    const double f1 = 1.0 + em;
    const double R = fast_rcp(den * f1);
    const double z = num * (R * f1);
    const double s = R * den;  // 1 / (1 + em)
    const double w = z * z;
    double p = 2.0 / 19.0;  // truncation: w^10 * 2/21 <= 2.4e-17 relative
    p = fma(p, w, 2.0 / 17.0);
    p = fma(p, w, 2.0 / 15.0);
    p = fma(p, w, 2.0 / 13.0);
    p = fma(p, w, 2.0 / 11.0);
    p = fma(p, w, 2.0 / 9.0);
    p = fma(p, w, 2.0 / 7.0);
    p = fma(p, w, 2.0 / 5.0);
    p = fma(p, w, 2.0 / 3.0);
    p = fma(p, w, 2.0);
    const double l1p = fma(z, p, big ? 0.69314718055994530942 : 0.0);
    const bool neg = t < 0.0;
    double lp_v = -((neg ? a : 0.0) + l1p);   // -log1p(exp(-t))
    double d_v = yd * (neg ? s : em * s);     // sign * e / (e + 1), e = exp(-t)
    if (t > 20.0) {
      lp_v = -em;
      d_v = -em;
    }
    if (t < -20.0) {
      lp_v = t;
      d_v = yd;
    }
    if (t != t) {
      lp_v = t;
      d_v = t;
    }
    lpi = lp_v;
    d = d_v;
  } else {
    const double xc = fmin(fmax(eta, -700.0), 700.0);
    double ee = fast_exp(xc);
    if (eta > 709.78) ee = INFINITY;
    if (eta != eta) ee = eta;
    lpi = yd * eta - ee;
    d = yd - ee;
  }
}

}  // namespace b200
