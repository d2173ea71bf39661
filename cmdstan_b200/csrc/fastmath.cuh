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
#include <cstring>

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

// ---- table-driven exp and the product form of the logistic log-likelihood (round 2) ---------------------------------
// On sm_100 every FP64 instruction of a link function is tensor (DMMA) or FMA time lost, so the per-observation cost is what
// counts.  Two reformulations take the logistic term from ~45 to 20 FP64 instructions without giving up accuracy:
//  * exp(x) = 2^k * 2^(j/32) * exp(r), |r| <= ln2/64: the 32-entry table (shared memory; one conflict-light LDS.64 on the
//    otherwise idle load pipe) shortens the polynomial from degree 12 to degree 6 (truncation 3.5e-18);
//  * sum_i log1p(e_i) = log prod_i (1 + e_i): the factors lie in (1, 2], so a running product needs ONE multiplication per term
//    instead of a reciprocal-scaled atanh series (13 instructions); its exponent is folded into an integer now and then
//    (integer pipe), and one log per accumulator is taken at the very end.  Error: each factor and each multiplication rounds
//    once (1.1e-16 relative to a product in [1, 2), i.e. absolute in the log), so a sum of n terms carries ~sqrt(n) 1.6e-16
//    absolute error -- relative 1e-15 or better whenever the terms themselves are not all below 1e-6;
//  * all comparisons (cut-offs +-20, the clamp at 700, NaN detection, the sign) are done on the bit patterns with integer
//    instructions, and the reciprocal takes one cubic correction step (3 FMA) after the hardware seed.
// The values reproduced are those of bernoulli_logit_glm_lpmf.hpp:121-126 / :137-142 and bernoulli_logit_lpmf.hpp:73-89.
#define B200_EXP2_TAB32                                                                                                          \
  0x1.0000000000000p+0, 0x1.059b0d3158574p+0, 0x1.0b5586cf9890fp+0, 0x1.11301d0125b51p+0, 0x1.172b83c7d517bp+0,                  \
  0x1.1d4873168b9aap+0, 0x1.2387a6e756238p+0, 0x1.29e9df51fdee1p+0, 0x1.306fe0a31b715p+0, 0x1.371a7373aa9cbp+0,                  \
  0x1.3dea64c123422p+0, 0x1.44e086061892dp+0, 0x1.4bfdad5362a27p+0, 0x1.5342b569d4f82p+0, 0x1.5ab07dd485429p+0,                  \
  0x1.6247eb03a5585p+0, 0x1.6a09e667f3bcdp+0, 0x1.71f75e8ec5f74p+0, 0x1.7a11473eb0187p+0, 0x1.82589994cce13p+0,                  \
  0x1.8ace5422aa0dbp+0, 0x1.93737b0cdc5e5p+0, 0x1.9c49182a3f090p+0, 0x1.a5503b23e255dp+0, 0x1.ae89f995ad3adp+0,                  \
  0x1.b7f76f2fb5e47p+0, 0x1.c199bdd85529cp+0, 0x1.cb720dcef9069p+0, 0x1.d5818dcfba487p+0, 0x1.dfc97337b9b5fp+0,                  \
  0x1.ea4afa2a490dap+0, 0x1.f50765b6e4540p+0
#ifdef B200_FASTMATH_HOST
static const double kExp2Tab32[32] = {B200_EXP2_TAB32};
#else
__device__ __constant__ const double kExp2Tab32[32] = {B200_EXP2_TAB32};  // 2^(j/32), correctly rounded (tools: mpmath)
#endif
// whole CTA: copy the table to shared memory (`tab`: 32 doubles); the caller synchronises before the first fast_exp_t
__device__ __forceinline__ void exp_table_load(double* tab, int tid) {
  if (tid < 32) tab[tid] = kExp2Tab32[tid];
}

// exp(x) for |x| <= 700 (callers clamp); tab = shared-memory copy of kExp2Tab32.  Relative error <= 3.5e-16.
__device__ __forceinline__ double fast_exp_t(double x, const double* tab) {
  const double kMagic = 6755399441055744.0;  // 1.5 * 2^52
  const double t = fma(x, 46.166241308446828384, kMagic);   // 32 / ln2
  const double nf = t - kMagic;
  const int n = __double2loint(t);                           // 32 k + j
  double r = fma(nf, -0x1.62e42fee00000p-6, x);              // ln2_hi / 32 (21 trailing zero bits: n * hi is exact)
  r = fma(nf, -5.9631716539705866e-12, r);                   // ln2_lo / 32
  double p = 1.0 / 720.0;
  p = fma(p, r, 1.0 / 120.0);
  p = fma(p, r, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const double v = tab[n & 31] * p;
  return __hiloint2double(__double2hiint(v) + ((n >> 5) << 20), __double2loint(v));
}

// 1/x for x in [1, 2]: hardware seed (>= 20 bits) + one cubic step r (1 + e + e^2), e = 1 - x r
__device__ __forceinline__ double fast_rcp3(double x) {
  double r;
#ifdef B200_FASTMATH_HOST
  {  // stands in for the hardware seed: 20 bits
    float f = 1.0f / (float)x;
    unsigned u;
    memcpy(&u, &f, 4);
    u &= 0xfffffff8u;
    memcpy(&f, &u, 4);
    r = (double)f;
  }
#else
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#endif
  const double e = fma(-x, r, 1.0);
  const double e1 = fma(e, e, e);
  return fma(r, e1, r);
}

// running sum of logistic log-likelihood terms: total = -(sum + expo ln2 + log(prod))
struct LogitAcc {
  double prod, sum;
  int expo;
};
__device__ __forceinline__ void logit_acc_init(LogitAcc& a) {
  a.prod = 1.0;
  a.sum = 0.0;
  a.expo = 0;
}
// move the exponent of the running product into the integer (call at least every ~900 terms)
__device__ __forceinline__ void logit_acc_fold(LogitAcc& a) {
  const int hi = __double2hiint(a.prod);
  a.expo += ((hi >> 20) & 0x7ff) - 1023;
  a.prod = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(a.prod));
}
// `pad_terms`: number of padding observations (eta == 0 exactly: zero rows / zero item slopes) that went through logit_term --
// each multiplied the product by exactly 2 and added nothing else, so they are taken out of the exponent, exactly
__device__ __forceinline__ double logit_acc_value(const LogitAcc& a, int pad_terms = 0) {
  return -(a.sum + ((a.expo - pad_terms) * 0.69314718055994530942 + log(a.prod)));
}

// One observation: adds its log-likelihood term to `acc` and returns d(term)/d(eta).  yd = 2y - 1 (exactly +-1).
//   t = yd eta;  term = -log1p(exp(-t)), t < -20: t, t > 20: -exp(-t);  derivative = yd e / (1 + e) with e = exp(-t),
//   t < -20: yd, t > 20: -e (the reference drops the sign there).
// Written for the instruction mix of sm_100 (one FP64 instruction occupies the pipe for two issue slots, so integer / select
// instructions cost as much as FP64 ones once they outnumber them): 19 FP64 + ~25 other instructions.
//  * with em := 0 below the lower cut-off, 1 + em == 1 and 1 / (1 + em) == 1 give term and derivative of that branch exactly;
//  * above the upper cut-off the product form differs from -e by e^2 / 2 < 3e-18; the derivative's sign quirk is a sign-bit edit;
//  * the derivative is yd (1/2 - copysign(s - 1/2, t)), s = 1 / (1 + em): no select, absolute error < 2.5e-16;
//  * a padding observation (eta == 0 exactly) needs no mask: see logit_acc_value;
//  * non-finite eta: +-inf behaves like the cut-off branches; NaN arises only from non-finite parameters, which the prior terms
//    of every model turn into a non-finite log density on their own (checked by the GPU edge-case tests).
// `negmask` carries the observation: 0 for y = 1, 0x80000000 for y = 0 (the sign bit of 2y - 1): multiplications by +-1 are sign-bit
// edits on the integer pipe.
__device__ __forceinline__ double logit_term(LogitAcc& acc, double eta, int negmask, const double* tab) {
  const int lo = __double2loint(eta);
  const int hi = __double2hiint(eta) ^ negmask;                  // t = (2y - 1) eta
  const unsigned ahi = (unsigned)hi & 0x7fffffffu;
  const bool big = ahi > 0x40340000u || (ahi == 0x40340000u && lo != 0);   // |t| > 20 (20.0 = 0x4034000000000000)
  const bool neg = hi < 0;
  double em = fast_exp_t(__hiloint2double((int)(ahi | 0x80000000u), lo), tab);   // exp(-|t|) in (0, 1]; garbage beyond 700, replaced next
  em = ((big && neg) || ahi >= 0x4085e000u) ? 0.0 : em;          // |t| >= 700 (0x4085e00000000000): exp(-|t|) < 1e-304 -> 0
  const double f1 = 1.0 + em;
  const double s = fast_rcp3(f1);                                // 1 / (1 + em) in [1/2, 1]
  const double h = s - 0.5;
  const double hs = __hiloint2double(__double2hiint(h) | (hi & 0x80000000), __double2loint(h));   // copysign(h, t), h >= 0
  const double w = 0.5 - hs;                                     // t < 0: s;  t >= 0: 1 - s = em s;  w >= 0
  const int dsign = (big && !neg) ? (int)0x80000000 : negmask;   // (2y - 1) w, except above the upper cut-off: -w
  const double d = __hiloint2double(__double2hiint(w) | dsign, __double2loint(w));
  acc.prod *= f1;
  const int m = hi >> 31;                                        // all ones when t < 0
  acc.sum += __hiloint2double((int)ahi & m, lo & m);             // max(-t, 0)
  return d;
}
// the same with the observation given as the double 2y - 1
__device__ __forceinline__ double logit_term(LogitAcc& acc, double eta, double yd, const double* tab) {
  return logit_term(acc, eta, __double2hiint(yd) & (int)0x80000000, tab);
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
