// Microbenchmark + layout check of the FP64 tensor-core shapes on sm_100a (mma.sync .f64: m8n8k4, m16n8k4, m16n8k8, m16n8k16).
//  (1) layout: D = A B with integer-valued A (16 x K), B (K x 8) loaded with the fragment layouts assumed below, against the host product;
//  (2) throughput: W warps per CTA, one CTA per SM, each warp a loop of NACC independent accumulators; reports TFLOP/s.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_shapes dmma_shapes.cu     Run: ./dmma_shapes
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

template <int SHAPE> struct Frag;   // number of doubles per lane of A, B, C and the K of the shape
template <> struct Frag<0> { static constexpr int A = 1, B = 1, C = 2, M = 8, K = 4; };    // m8n8k4
template <> struct Frag<1> { static constexpr int A = 2, B = 1, C = 4, M = 16, K = 4; };   // m16n8k4
template <> struct Frag<2> { static constexpr int A = 4, B = 2, C = 4, M = 16, K = 8; };   // m16n8k8
template <> struct Frag<3> { static constexpr int A = 8, B = 4, C = 4, M = 16, K = 16; };  // m16n8k16

template <int SHAPE>
__device__ __forceinline__ void mma(double (&c)[Frag<SHAPE>::C], const double (&a)[Frag<SHAPE>::A], const double (&b)[Frag<SHAPE>::B]) {
  if constexpr (SHAPE == 0) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a[0]), "d"(b[0]));
  } else if constexpr (SHAPE == 1) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
  } else if constexpr (SHAPE == 2) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
  } else {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
  }
}

// assumed layouts (PTX ISA, "matrix fragments for mma with .f64"): g = lane >> 2, t = lane & 3
//   A (M x K, row): a_i -> row g + 8 (i & 1) [M = 16] / g [M = 8], col t + 4 (i >> 1) [M = 16] / t [M = 8]
//   B (K x 8, col): b_i -> row t + 4 i, col g
//   C (M x 8)     : M = 8: c_i -> row g, col 2 t + i;  M = 16: c_i -> row g + 8 (i >> 1), col 2 t + (i & 1)
template <int SHAPE>
__global__ void layout_kernel(const double* A, const double* B, double* D) {
  using F = Frag<SHAPE>;
  const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
  double a[F::A], b[F::B], c[F::C];
  for (int i = 0; i < F::A; i++) {
    const int r = F::M == 16 ? g + 8 * (i & 1) : g, col = F::M == 16 ? t + 4 * (i >> 1) : t;
    a[i] = A[r * F::K + col];
  }
  for (int i = 0; i < F::B; i++) b[i] = B[(t + 4 * i) * 8 + g];
  for (int i = 0; i < F::C; i++) c[i] = 0;
  mma<SHAPE>(c, a, b);
  for (int i = 0; i < F::C; i++) {
    const int r = F::M == 16 ? g + 8 * (i >> 1) : g, col = F::M == 16 ? 2 * t + (i & 1) : 2 * t + i;
    D[r * 8 + col] = c[i];
  }
}

template <int SHAPE, int NACC>
__global__ void __launch_bounds__(512, 1) rate_kernel(double* out, int iters) {
  using F = Frag<SHAPE>;
  double a[F::A], b[F::B], c[NACC][F::C];
  for (int i = 0; i < F::A; i++) a[i] = 1.0 + threadIdx.x * 1e-9 + i;
  for (int i = 0; i < F::B; i++) b[i] = 1e-3 * (i + 1);
  for (int n = 0; n < NACC; n++)
    for (int i = 0; i < F::C; i++) c[n][i] = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int n = 0; n < NACC; n++) mma<SHAPE>(c[n], a, b);
  }
  double s = 0;
  for (int n = 0; n < NACC; n++)
    for (int i = 0; i < F::C; i++) s += c[n][i];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int SHAPE>
bool check_layout() {
  using F = Frag<SHAPE>;
  std::vector<double> A(F::M * F::K), B(F::K * 8), D(F::M * 8, -1), R(F::M * 8, 0);
  for (int i = 0; i < F::M * F::K; i++) A[i] = (double)((i * 7 + 3) % 11) - 5;
  for (int i = 0; i < F::K * 8; i++) B[i] = (double)((i * 5 + 1) % 13) - 6;
  for (int r = 0; r < F::M; r++)
    for (int c = 0; c < 8; c++)
      for (int k = 0; k < F::K; k++) R[r * 8 + c] += A[r * F::K + k] * B[k * 8 + c];
  double *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 8); cudaMalloc(&dB, B.size() * 8); cudaMalloc(&dD, D.size() * 8);
  cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice);
  layout_kernel<SHAPE><<<1, 32>>>(dA, dB, dD);
  cudaMemcpy(D.data(), dD, D.size() * 8, cudaMemcpyDeviceToHost);
  bool ok = cudaGetLastError() == cudaSuccess;
  for (size_t i = 0; i < D.size(); i++) ok = ok && D[i] == R[i];
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return ok;
}

template <int SHAPE, int NACC>
double rate(int warps, int sms) {
  using F = Frag<SHAPE>;
  const int iters = 20000 / (F::K / 4) / (F::M / 8);
  double* out;
  cudaMalloc(&out, sizeof(double) * sms * warps * 32);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    rate_kernel<SHAPE, NACC><<<sms, warps * 32>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fl = 2.0 * F::M * 8 * F::K * (double)NACC * iters * warps * sms;
    if (rep) best = fl / (ms * 1e-3) / 1e12 > best ? fl / (ms * 1e-3) / 1e12 : best;
  }
  cudaFree(out);
  return best;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  std::printf("layout m8n8k4 %s, m16n8k4 %s, m16n8k8 %s, m16n8k16 %s\n", check_layout<0>() ? "ok" : "WRONG", check_layout<1>() ? "ok" : "WRONG",
              check_layout<2>() ? "ok" : "WRONG", check_layout<3>() ? "ok" : "WRONG");
  std::printf("TFLOP/s (fp64) on %d SMs, one CTA per SM, 8 independent accumulators per warp\n", sms);
  std::printf("%-10s %8s %8s %8s %8s\n", "warps/CTA", "m8n8k4", "m16n8k4", "m16n8k8", "m16n8k16");
  for (int w : {4, 8, 12, 16})
    std::printf("%-10d %8.2f %8.2f %8.2f %8.2f\n", w, rate<0, 8>(w, sms), rate<1, 8>(w, sms), rate<2, 8>(w, sms), rate<3, 8>(w, sms));
  std::printf("2 independent accumulators per warp\n");
  for (int w : {4, 8, 16})
    std::printf("%-10d %8.2f %8.2f %8.2f %8.2f\n", w, rate<0, 2>(w, sms), rate<1, 2>(w, sms), rate<2, 2>(w, sms), rate<3, 2>(w, sms));
  return 0;
}
