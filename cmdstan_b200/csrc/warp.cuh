// Warp-cooperation primitives for the chain kernels: one NUTS chain occupies one warp, lane l owns vector
// elements l, l+32, l+64, ... of every per-chain vector ("owner lane" rule: an element is only ever read or
// written by its owner, so no __syncwarp is needed between vector passes).  Scalars are warp-uniform: every
// lane holds the same value, reductions are xor-butterflies (bitwise identical result in all lanes, fixed
// summation order => run-to-run reproducible).
//
// B200_HOST_EMUL builds the same source for the host with a "warp" of one lane; it exists only so that the
// CPU test-suite can drive the device state machine without a GPU (tests/host_emul).  It is never part of the
// shipped library.
#pragma once
#include <cstdint>
#include <cstring>

#ifdef B200_HOST_EMUL
#include <cmath>
#define B200_DEV inline
#define B200_DEV_CALL inline
#define B200_HD inline
namespace b200 {
constexpr int kWarp = 1;
constexpr unsigned kFull = 1u;
inline int lane_id() { return 0; }
inline double bcast0(double v) { return v; }
inline int bcast0i(int v) { return v; }
inline double warp_sum(double v) { return v; }
inline void warp_sum6(double&, double&, double&, double&, double&, double&) {}
inline void warp_sync() {}
inline double dev_exp(double x) { return std::exp(x); }
inline double dev_log(double x) { return std::log(x); }
inline double dev_log1p(double x) { return std::log1p(x); }
}  // namespace b200
#else
#define B200_DEV __device__ __forceinline__
// out-of-line device functions: the chain kernels are branchy state machines whose warps sit at different program
// points, so code size (instruction-cache footprint) matters more than call overhead
#define B200_DEV_CALL static __device__ __noinline__
#define B200_HD __host__ __device__ inline
namespace b200 {
constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;
B200_DEV int lane_id() { return threadIdx.x & 31; }
B200_DEV double bcast0(double v) { return __shfl_sync(kFull, v, 0); }
B200_DEV int bcast0i(int v) { return __shfl_sync(kFull, v, 0); }
B200_DEV_CALL double warp_sum(double v) {  // out of line: ~50 call sites in the chain kernels (instruction-cache budget)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
// six sums at once: the same xor butterfly per value (bit-identical to six warp_sum calls), but the 30 shuffles form 5
// dependent rounds of 6 independent ones instead of 30 dependent steps -- the U-turn criteria of every tree merge need six dot
// products, and the chain kernels are latency-bound.  Inlined (two call sites): reference parameters of an out-of-line
// function would live in local memory
B200_DEV void warp_sum6(double& a, double& b, double& c, double& d, double& e, double& f) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ta = __shfl_xor_sync(kFull, a, o), tb = __shfl_xor_sync(kFull, b, o), tc = __shfl_xor_sync(kFull, c, o);
    const double td = __shfl_xor_sync(kFull, d, o), te = __shfl_xor_sync(kFull, e, o), tf = __shfl_xor_sync(kFull, f, o);
    a += ta; b += tb; c += tc; d += td; e += te; f += tf;
  }
}
B200_DEV void warp_sync() { __syncwarp(); }
// one out-of-line copy of each double-precision transcendental: the chain kernels are instruction-cache bound
// (28 warps per SM sitting at different points of a branchy state machine), so code size is the budget
B200_DEV_CALL double dev_exp(double x) { return exp(x); }
B200_DEV_CALL double dev_log(double x) { return log(x); }
B200_DEV_CALL double dev_log1p(double x) { return log1p(x); }
}  // namespace b200
#endif
