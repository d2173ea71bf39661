// Per-chain random streams, bit-compatible with the reference's stan::rng_t = boost::random::mixmax
// (mixmax_engine<17,36,0>) as created by stan::services::util::create_rng(seed, chain)
// (S/services/util/create_rng.hpp:28-34  ->  mixmax(clusterID=0, machineID=1, runID=seed, streamID=chain)).
//
//   engine step / refill     B/random/mixmax.hpp:104-112 (operator()), :194-208 (iterate_raw_vec)
//   stream seeding           B/random/mixmax.hpp:220-286 (seed_uniquestream / apply_bigskip), :288-301 (fmodmulM61)
//   uniform_01               B/random/uniform_01.hpp:137-156      (u * 2^-61, redraw on >= 1)
//   uniform_real(a,b)        B/random/uniform_real_distribution.hpp:47-64
//   normal (ziggurat, 128)   B/random/normal_distribution.hpp:120-201, detail/int_float_pair.hpp:34-98
//   exponential (tail, 256)  B/random/exponential_distribution.hpp:181-218
//
// The skip-coefficient table and the ziggurat layer tables are *derived* (tools/gen_mixmax_skip.py,
// tools/gen_ziggurat_tables.py), not copied; tests/test_tables.py checks them against the reference headers.
//
// Layout: 19 x u64 per chain = V[17], sumtot, counter.  The stream is consumed by lane 0 of the chain's warp and
// the variate is broadcast with a shuffle, so all control flow that depends on random numbers stays warp-uniform.
#pragma once
#include "warp.cuh"

namespace b200 {

constexpr int kRngWords = 19;
constexpr unsigned long long kM61 = 2305843009213693951ULL;

#ifdef B200_HOST_EMUL
#define B200_TABLE static const
#else
#define B200_TABLE __device__ const
#endif
namespace tables {
#include "ziggurat_tables.inc"
}  // namespace tables

B200_HD unsigned long long mm_mod(unsigned long long k) { return (k & kM61) + (k >> 61); }
B200_HD unsigned long long mm_mulwu(unsigned long long k) { return ((k << 36) & kM61) ^ (k >> 25); }

// one application of the MIXMAX matrix to the state vector (refill of 16 outputs)
B200_HD unsigned long long mm_iterate(unsigned long long* Y, unsigned long long sum_old) {
  unsigned long long tempP = 0, tempV = sum_old;
  Y[0] = tempV;
  unsigned long long sumtot = Y[0], ovflow = 0;
  for (int i = 1; i < 17; i++) {
    unsigned long long tempPO = mm_mulwu(tempP);
    tempV = tempV + tempPO;
    tempP = mm_mod(tempP + Y[i]);
    tempV = mm_mod(tempV + tempP);
    Y[i] = tempV;
    sumtot += tempV;
    if (sumtot < tempV) ovflow++;
  }
  return mm_mod(mm_mod(sumtot) + (ovflow << 3));
}

// ---- single-lane primitives (call from lane 0 only) ---------------------------------------------------------
B200_DEV_CALL unsigned long long rng_next(unsigned long long* st) {
  int counter = (int)st[18];
  if (counter <= 16) {
    st[18] = (unsigned long long)(counter + 1);
    return st[counter];
  }
  st[17] = mm_iterate(st, st[17]);
  st[18] = 2;
  return st[1];
}

B200_DEV double rng_uniform01(unsigned long long* st) {
  for (;;) {
    double r = (double)rng_next(st) * 4.336808689942017736029811203479766845703125e-19;  // 2^-61
    if (r < 1.0) return r;
  }
}

B200_DEV double rng_uniform_real(unsigned long long* st, double a, double b) {
  for (;;) {
    double num = (double)rng_next(st);
    double r = num / 2305843009213693952.0 * (b - a) + a;
    if (r < b) return r;
  }
}

B200_DEV_CALL double rng_unit_exponential(unsigned long long* st) {
  using namespace tables;
  double shift = 0;
  for (;;) {
    unsigned long long u = rng_next(st);
    int i = (int)(u & 0xFF);
    double x = (double)(u >> 8) * 1.1102230246251565404236316680908203125e-16 * kZigExpX[i];  // 2^-53
    if (x < kZigExpX[i + 1]) return shift + x;
    if (i == 0) {
      shift += kZigExpX[1];
    } else {
      double y01 = rng_uniform01(st);
      double y = kZigExpY[i] + y01 * (kZigExpY[i + 1] - kZigExpY[i]);
      double y_above_ubound = (kZigExpX[i] - kZigExpX[i + 1]) * y01 - (kZigExpX[i] - x);
      double y_above_lbound = y - (kZigExpY[i + 1] + (kZigExpX[i + 1] - x) * kZigExpY[i + 1]);
      if (y_above_ubound < 0 && (y_above_lbound < 0 || y < dev_exp(-x))) return x + shift;
    }
  }
}

B200_DEV_CALL double rng_normal(unsigned long long* st) {
  using namespace tables;
  for (;;) {
    unsigned long long u = rng_next(st);
    int bucket = (int)(u & 0xFF);
    double r = (double)(u >> 8) * 1.1102230246251565404236316680908203125e-16;
    int sign = (bucket & 1) * 2 - 1;
    int i = bucket >> 1;
    double x = r * kZigNormX[i];
    if (x < kZigNormX[i + 1]) return x * sign;
    if (i == 0) {
      const double tail_start = kZigNormX[1];
      for (;;) {
        double tx = rng_unit_exponential(st) / tail_start;
        double ty = rng_unit_exponential(st);
        if (2 * ty > tx * tx) return (tx + tail_start) * sign;
      }
    }
    double y01 = rng_uniform01(st);
    double y = kZigNormY[i] + y01 * (kZigNormY[i + 1] - kZigNormY[i]);
    double y_above_ubound, y_above_lbound;
    if (kZigNormX[i] >= 1) {
      y_above_ubound = (kZigNormX[i] - kZigNormX[i + 1]) * y01 - (kZigNormX[i] - x);
      y_above_lbound = y - (kZigNormY[i] + (kZigNormX[i] - x) * kZigNormY[i] * kZigNormX[i]);
    } else {
      y_above_lbound = (kZigNormX[i] - kZigNormX[i + 1]) * y01 - (kZigNormX[i] - x);
      y_above_ubound = y - (kZigNormY[i] + (kZigNormX[i] - x) * kZigNormY[i] * kZigNormX[i]);
    }
    if (y_above_ubound < 0 && (y_above_lbound < 0 || y < dev_exp(-(x * x / 2)))) return x * sign;
  }
}

// ---- warp-level wrappers: lane 0 draws, everybody gets the value ---------------------------------------------
B200_DEV double warp_uniform01(unsigned long long* st, int lane) {
  double v = 0.0;
  if (lane == 0) v = rng_uniform01(st);
  return bcast0(v);
}

// fills dst[0..n) with variates in stream order; element i is stored by its owner lane (i % warp)
// kind 0: normal(0,1) scaled by 1/sqrt(scale[i])  (diag_e_metric::sample_p, S/mcmc/hmc/hamiltonians/diag_e_metric.hpp:44-50);
//         scale == nullptr: unscaled (the u of dense_e_metric::sample_p)
// kind 1: uniform_real(-radius, radius)           (random_var_context, S/io/random_var_context.hpp:53-56)
B200_DEV_CALL void warp_fill(unsigned long long* st, int lane, double* dst, int n, int kind, const double* scale, double radius) {
  for (int base = 0; base < n; base += kWarp) {
    int m = n - base < kWarp ? n - base : kWarp;
    double mine = 0.0;
    for (int j = 0; j < m; j++) {
      double v = 0.0;
      if (lane == 0) v = kind == 0 ? rng_normal(st) : rng_uniform_real(st, -radius, radius);
      v = bcast0(v);
      if (lane == j) mine = v;
    }
    if (lane < m) dst[base + lane] = (kind == 0 && scale) ? mine / sqrt(scale[base + lane]) : mine;
  }
}

}  // namespace b200
