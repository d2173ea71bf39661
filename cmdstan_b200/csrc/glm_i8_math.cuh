// The arithmetic of the INT8 digit-plane GLM kernel (glm_i8.cu) that does not touch the tensor core: fixed-point format, digit
// extraction, recombination of the INT32 diagonals, the staged link.  A header of its own so that the CPU suite can build it for the
// host (tests/host_emul/i8_math_host.cpp, tests/test_i8_math_host.py) and compare it bit for bit with the numpy restatement of the
// kernel's numerics (tools/ozaki_i8_model.py) and with fastmath.cuh's logit_term -- the kernel itself needs tcgen05 and runs on
// hardware only.
#pragma once
#include <cstdint>

#include "fastmath.cuh"

namespace b200 {

template <int NS> struct I8Fmt {
  static constexpr int F = 8 * NS - 2;                      // fixed-point fraction bits of every operand
  static constexpr int SEXP = -2 * F + 8 * (NS - 1);         // sum_d v w = S 2^SEXP, S = sum_k diag_k 256^(NS-1-k)
};

// Balanced base-256 digits of V = rint(v 2^F) for four values at once.  Adding 0x80 at each of the NS - 1 lower byte positions
// turns the digits into the bytes of V + BIAS (each byte = digit + 128, carries propagate as they must), and BIAS is folded into
// the magic constant of the double -> integer conversion: the mantissa of fma(v, 2^F, 1.5 2^52 + BIAS) holds 2^51 + V + BIAS.  One
// FP64 instruction per value; the "+ 128" is undone after the byte transposes, once per packed word of four chains.
template <int NS>
__device__ __forceinline__ void i8_digit_planes4(const double (&v)[4], uint32_t (&pk)[NS]) {
  constexpr double kMagic = 6755399441055744.0 + (NS == 6 ? 551911719040.0 /* 0x8080808080 */ : 2155905152.0 /* 0x80808080 */);
  uint32_t lo[4], hi[4];
#pragma unroll
  for (int c = 0; c < 4; c++) {
    const double t = fma(v[c], (double)(1ll << I8Fmt<NS>::F), kMagic);
    lo[c] = (uint32_t)__double2loint(t);
    hi[c] = (uint32_t)__double2hiint(t);
  }
  // byte b of lo is digit plane NS - 1 - b; hi: plane 0 in byte 0 (NS = 5); plane 1 in byte 0 and plane 0 in byte 1 (NS = 6)
  const uint32_t l01a = __byte_perm(lo[0], lo[1], 0x5140), l01b = __byte_perm(lo[0], lo[1], 0x7362);   // (b0 b0' b1 b1'), (b2 b2' b3 b3')
  const uint32_t l23a = __byte_perm(lo[2], lo[3], 0x5140), l23b = __byte_perm(lo[2], lo[3], 0x7362);
  pk[NS - 1] = __byte_perm(l01a, l23a, 0x5410) ^ 0x80808080u;
  pk[NS - 2] = __byte_perm(l01a, l23a, 0x7632) ^ 0x80808080u;
  pk[NS - 3] = __byte_perm(l01b, l23b, 0x5410) ^ 0x80808080u;
  pk[NS - 4] = __byte_perm(l01b, l23b, 0x7632) ^ 0x80808080u;
  const uint32_t h01 = __byte_perm(hi[0], hi[1], 0x5140), h23 = __byte_perm(hi[2], hi[3], 0x5140);
  if (NS == 6) {
    pk[1] = __byte_perm(h01, h23, 0x5410) ^ 0x80808080u;
    pk[0] = __byte_perm(h01, h23, 0x7632);
  } else {
    pk[0] = __byte_perm(h01, h23, 0x5410);
  }
}

// fastmath.cuh's logit_term for four chains at once, written stage by stage: the FP64 instructions of one chain form a dependent
// chain of ~20 (magic-number rounding, argument reduction, degree-6 polynomial, table factor, reciprocal with one cubic step), each
// waiting for its predecessor; stated per chain, the compiler (96 registers) issues them chain after chain and a warp spends ~23
// cycles per FP64 instruction (r02q / r02t: 6 900 cycles of link per tile).  Four independent chains per stage fill those gaps.
// Same operations in the same order per chain as logit_term: same bits.  prod / sum: the running product of the factors 1 + e and
// the running sum of max(-t, 0) (LogitAcc without its exponent: at most ~900 factors in (1, 2] are multiplied between two folds).
__device__ __forceinline__ void logit_terms4(double (&prod)[4], double (&sum)[4], const double (&eta)[4], int negmask, const double* tab, double (&d)[4]) {
  const double kMagic = 6755399441055744.0;   // 1.5 * 2^52
  int lo[4], hi[4], n[4];
  unsigned ahi[4];
  double x[4], t[4], nf[4], r[4], p[4], tv[4], em[4], f1[4], s[4], e[4];
#pragma unroll
  for (int c = 0; c < 4; c++) {
    lo[c] = __double2loint(eta[c]);
    hi[c] = __double2hiint(eta[c]) ^ negmask;                       // t = (2y - 1) eta
    ahi[c] = (unsigned)hi[c] & 0x7fffffffu;
    x[c] = __hiloint2double((int)(ahi[c] | 0x80000000u), lo[c]);    // -|t|
  }
#pragma unroll
  for (int c = 0; c < 4; c++) t[c] = fma(x[c], 46.166241308446828384, kMagic);   // 32 / ln2
#pragma unroll
  for (int c = 0; c < 4; c++) {
    nf[c] = t[c] - kMagic;
    n[c] = __double2loint(t[c]);
  }
#pragma unroll
  for (int c = 0; c < 4; c++) r[c] = fma(nf[c], -0x1.62e42fee00000p-6, x[c]);
#pragma unroll
  for (int c = 0; c < 4; c++) r[c] = fma(nf[c], -5.9631716539705866e-12, r[c]);
#pragma unroll
  for (int c = 0; c < 4; c++) {
    tv[c] = tab[n[c] & 31];
    p[c] = fma(1.0 / 720.0, r[c], 1.0 / 120.0);
  }
#pragma unroll
  for (int c = 0; c < 4; c++) p[c] = fma(p[c], r[c], 1.0 / 24.0);
#pragma unroll
  for (int c = 0; c < 4; c++) p[c] = fma(p[c], r[c], 1.0 / 6.0);
#pragma unroll
  for (int c = 0; c < 4; c++) p[c] = fma(p[c], r[c], 0.5);
#pragma unroll
  for (int c = 0; c < 4; c++) p[c] = fma(p[c], r[c], 1.0);
#pragma unroll
  for (int c = 0; c < 4; c++) p[c] = fma(p[c], r[c], 1.0);
#pragma unroll
  for (int c = 0; c < 4; c++) {
    const double v = tv[c] * p[c];
    em[c] = __hiloint2double(__double2hiint(v) + ((n[c] >> 5) << 20), __double2loint(v));   // exp(-|t|)
    const bool big = ahi[c] > 0x40340000u || (ahi[c] == 0x40340000u && lo[c] != 0);          // |t| > 20
    em[c] = ((big && hi[c] < 0) || ahi[c] >= 0x4085e000u) ? 0.0 : em[c];
  }
#pragma unroll
  for (int c = 0; c < 4; c++) f1[c] = 1.0 + em[c];
#pragma unroll
  for (int c = 0; c < 4; c++) {
#ifdef B200_FASTMATH_HOST
    {  // the stand-in of fast_rcp3 for the hardware seed (20 bits)
      float f = 1.0f / (float)f1[c];
      unsigned u;
      memcpy(&u, &f, 4);
      u &= 0xfffffff8u;
      memcpy(&f, &u, 4);
      s[c] = (double)f;
    }
#else
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(s[c]) : "d"(f1[c]));
#endif
  }
#pragma unroll
  for (int c = 0; c < 4; c++) e[c] = fma(-f1[c], s[c], 1.0);
#pragma unroll
  for (int c = 0; c < 4; c++) e[c] = fma(e[c], e[c], e[c]);
#pragma unroll
  for (int c = 0; c < 4; c++) s[c] = fma(s[c], e[c], s[c]);                                    // 1 / (1 + em)
#pragma unroll
  for (int c = 0; c < 4; c++) {
    const double h = s[c] - 0.5;
    const double hs = __hiloint2double(__double2hiint(h) | (hi[c] & (int)0x80000000), __double2loint(h));   // copysign(h, t)
    const double w = 0.5 - hs;
    const bool big = ahi[c] > 0x40340000u || (ahi[c] == 0x40340000u && lo[c] != 0);
    const int dsign = (big && hi[c] >= 0) ? (int)0x80000000 : negmask;
    d[c] = __hiloint2double(__double2hiint(w) | dsign, __double2loint(w));
    prod[c] *= f1[c];
    const int m = hi[c] >> 31;
    sum[c] += __hiloint2double((int)ahi[c] & m, lo[c] & m);                                     // max(-t, 0)
  }
}

// S = sum_k d[k] 256^(NS-1-k) from the INT32 diagonals: three (NS = 6) wide multiply-adds and two 64-bit additions instead of a
// Horner chain of 64-bit shifts
template <int NS>
__device__ __forceinline__ long long i8_combine(const int (&d)[NS]) {
  // the upper diagonals are small (the top digits are at most 64): d0 256 + d1 and d2 256 + d3 fit 32 bits for K <= 128
  if (NS == 6) {
    const int u3 = d[0] * 256 + d[1], u2 = d[2] * 256 + d[3];
    const long long w1 = (long long)d[4] * 256 + d[5];
    return ((long long)u3 << 32) + ((long long)u2 << 16) + w1;
  } else {
    const int u2 = d[1] * 256 + d[2];
    const long long w1 = (long long)d[3] * 256 + d[4];
    return ((long long)d[0] << 32) + ((long long)u2 << 16) + w1;
  }
}

}  // namespace b200
