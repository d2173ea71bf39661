// TEST-ONLY: builds the tensor-core-free arithmetic of the INT8 digit-plane GLM kernel (cmdstan_b200/csrc/glm_i8_math.cuh) for the
// host, with bit-exact stand-ins for the device intrinsics, so that the CPU suite can compare it with the numpy restatement of the
// kernel's numerics (tools/ozaki_i8_model.py) and with fastmath.cuh's logit_term (tests/test_i8_math_host.py).
#include <cmath>
#include <cstdint>
#include <cstring>

#define __device__
#define __forceinline__ inline
static inline int __double2loint(double x) { uint64_t b; std::memcpy(&b, &x, 8); return (int)(uint32_t)(b & 0xffffffffu); }
static inline int __double2hiint(double x) { uint64_t b; std::memcpy(&b, &x, 8); return (int)(uint32_t)(b >> 32); }
static inline double __hiloint2double(int hi, int lo) { uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double x; std::memcpy(&x, &b, 8); return x; }
// PRMT in its default mode: byte i of the result is byte (nibble i of the selector, 0 .. 7) of the pair {y, x}
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s) {
  const uint64_t src = ((uint64_t)y << 32) | x;
  uint32_t r = 0;
  for (int i = 0; i < 4; i++) r |= (uint32_t)((src >> (8 * ((s >> (4 * i)) & 7))) & 0xff) << (8 * i);
  return r;
}
using std::fma;
#define B200_FASTMATH_HOST 1
#include "../../cmdstan_b200/csrc/glm_i8_math.cuh"

// digit planes of n values (n % 4 == 0): planes[s][i] = digit s of value i as a signed byte
template <int NS>
static void planes(const double* v, int n, signed char* out) {
  for (int i = 0; i < n; i += 4) {
    const double v4[4] = {v[i], v[i + 1], v[i + 2], v[i + 3]};
    uint32_t pk[NS];
    b200::i8_digit_planes4<NS>(v4, pk);
    for (int s = 0; s < NS; s++)
      for (int c = 0; c < 4; c++) out[(size_t)s * n + i + c] = (signed char)((pk[s] >> (8 * c)) & 0xff);
  }
}
extern "C" void i8_planes(int ns, const double* v, int n, signed char* out) {
  if (ns == 6) planes<6>(v, n, out);
  else planes<5>(v, n, out);
}
extern "C" void i8_combine(int ns, const int* diag, int n, long long* out) {   // diag [n][ns]
  for (int i = 0; i < n; i++) {
    if (ns == 6) {
      int d[6];
      for (int k = 0; k < 6; k++) d[k] = diag[i * 6 + k];
      out[i] = b200::i8_combine<6>(d);
    } else {
      int d[5];
      for (int k = 0; k < 5; k++) d[k] = diag[i * 5 + k];
      out[i] = b200::i8_combine<5>(d);
    }
  }
}
// the staged link against logit_term, element by element: returns the number of elements whose derivative, running product or
// running sum differ in any bit (the reciprocal seed is the same stand-in in both: B200_FASTMATH_HOST)
extern "C" int i8_link_vs_logit_term(const double* eta, const int* y, int n, double* d_out) {
  int bad = 0;
  for (int i = 0; i + 4 <= n; i += 4) {
    for (int yy = 0; yy < 2; yy++) {
      const int negmask = yy ? 0 : (int)0x80000000;
      double prod[4] = {1.25, 1.0, 1.5, 1.0}, sum[4] = {0.0, 3.0, 0.5, 0.0}, e4[4], d4[4];
      for (int c = 0; c < 4; c++) e4[c] = eta[i + c];
      b200::logit_terms4(prod, sum, e4, negmask, b200::kExp2Tab32, d4);
      for (int c = 0; c < 4; c++) {
        b200::LogitAcc a;
        a.prod = (const double[]){1.25, 1.0, 1.5, 1.0}[c];
        a.sum = (const double[]){0.0, 3.0, 0.5, 0.0}[c];
        a.expo = 0;
        const double d = b200::logit_term(a, eta[i + c], negmask, b200::kExp2Tab32);
        if (std::memcmp(&d, &d4[c], 8) || std::memcmp(&a.prod, &prod[c], 8) || std::memcmp(&a.sum, &sum[c], 8)) bad++;
        if (yy == (y[i + c] != 0)) d_out[i + c] = d4[c];
      }
    }
  }
  return bad;
}
