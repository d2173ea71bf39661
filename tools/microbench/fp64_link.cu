// What bounds the FP64 link of the INT8 GLM kernel?  (1) DFMA latency / throughput per SM sub-partition as a function of the
// independent chains per warp and the warps per sub-partition; (2) the kernel's own per-element code (recombination of six INT32
// diagonals, logit_terms4, digit extraction) on register data, 4 / 8 / 16 warps per SM sub-partition.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../cmdstan_b200/csrc -o fp64_link fp64_link.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "fastmath.cuh"
using namespace b200;

template <int ILP>
__global__ void dfma_kernel(double* out, int iters, double a, double b, long long* cyc) {
  double x[ILP];
  for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-3 + i;
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++)
#pragma unroll
      for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
  }
  const long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < ILP; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// the staged link of glm_i8.cu (copied: this file is a measurement, the product code is csrc/glm_i8.cu)
__device__ __forceinline__ void logit_terms4(double (&prod)[4], double (&sum)[4], const double (&eta)[4], int negmask, const double* tab, double (&d)[4]) {
  const double kMagic = 6755399441055744.0;
  int lo[4], hi[4], n[4];
  unsigned ahi[4];
  double x[4], t[4], nf[4], r[4], p[4], tv[4], em[4], f1[4], s[4], e[4];
#pragma unroll
  for (int c = 0; c < 4; c++) { lo[c] = __double2loint(eta[c]); hi[c] = __double2hiint(eta[c]) ^ negmask; ahi[c] = (unsigned)hi[c] & 0x7fffffffu; x[c] = __hiloint2double((int)(ahi[c] | 0x80000000u), lo[c]); }
#pragma unroll
  for (int c = 0; c < 4; c++) t[c] = fma(x[c], 46.166241308446828384, kMagic);
#pragma unroll
  for (int c = 0; c < 4; c++) { nf[c] = t[c] - kMagic; n[c] = __double2loint(t[c]); }
#pragma unroll
  for (int c = 0; c < 4; c++) r[c] = fma(nf[c], -0x1.62e42fee00000p-6, x[c]);
#pragma unroll
  for (int c = 0; c < 4; c++) r[c] = fma(nf[c], -5.9631716539705866e-12, r[c]);
#pragma unroll
  for (int c = 0; c < 4; c++) { tv[c] = tab[n[c] & 31]; p[c] = fma(1.0 / 720.0, r[c], 1.0 / 120.0); }
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
    em[c] = __hiloint2double(__double2hiint(v) + ((n[c] >> 5) << 20), __double2loint(v));
    const bool big = ahi[c] > 0x40340000u || (ahi[c] == 0x40340000u && lo[c] != 0);
    em[c] = ((big && hi[c] < 0) || ahi[c] >= 0x4085e000u) ? 0.0 : em[c];
  }
#pragma unroll
  for (int c = 0; c < 4; c++) f1[c] = 1.0 + em[c];
#pragma unroll
  for (int c = 0; c < 4; c++) asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(s[c]) : "d"(f1[c]));
#pragma unroll
  for (int c = 0; c < 4; c++) e[c] = fma(-f1[c], s[c], 1.0);
#pragma unroll
  for (int c = 0; c < 4; c++) e[c] = fma(e[c], e[c], e[c]);
#pragma unroll
  for (int c = 0; c < 4; c++) s[c] = fma(s[c], e[c], s[c]);
#pragma unroll
  for (int c = 0; c < 4; c++) {
    const double h = s[c] - 0.5;
    const double hs = __hiloint2double(__double2hiint(h) | (hi[c] & (int)0x80000000), __double2loint(h));
    const double w = 0.5 - hs;
    const bool big = ahi[c] > 0x40340000u || (ahi[c] == 0x40340000u && lo[c] != 0);
    const int dsign = (big && hi[c] >= 0) ? (int)0x80000000 : negmask;
    d[c] = __hiloint2double(__double2hiint(w) | dsign, __double2loint(w));
    prod[c] *= f1[c];
    const int m = hi[c] >> 31;
    sum[c] += __hiloint2double((int)ahi[c] & m, lo[c] & m);
  }
}
__device__ __forceinline__ void digits6(double v, uint32_t& lo, uint32_t& hi) {
  const double t = fma(v, (double)(1ll << 46), 6755399441055744.0);
  const uint32_t vlo = (uint32_t)__double2loint(t);
  const uint32_t vhi = ((uint32_t)__double2hiint(t) & 0xfffffu) - 0x80000u;
  lo = vlo + 0x80808080u;
  const uint32_t carry = lo < vlo ? 1u : 0u;
  hi = (vhi + 0x80u + carry) ^ 0x80u;
  lo ^= 0x80808080u;
}

// MODE 0: link only; 1: + recombination of six diagonals; 2: + digits and byte transposes (the whole per-element work)
template <int MODE>
__global__ void link_kernel(const int* diag, double* out, uint32_t* outw, int iters, long long* cyc) {
  __shared__ double tab[32];
  if (threadIdx.x < 32) tab[threadIdx.x] = kExp2Tab32[threadIdx.x];
  __syncthreads();
  double prod[4] = {1, 1, 1, 1}, sum[4] = {0, 0, 0, 0};
  int v[6][4];
  for (int k = 0; k < 6; k++)
    for (int c = 0; c < 4; c++) v[k][c] = diag[(threadIdx.x * 6 + k) * 4 + c];
  uint32_t acc = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    double eta[4], d[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
      if (MODE >= 1) {
        const long long w3 = (long long)v[0][c] * 256 + v[1][c], w2 = (long long)v[2][c] * 256 + v[3][c], w1 = (long long)v[4][c] * 256 + v[5][c];
        eta[c] = (double)((w3 << 32) + (w2 << 16) + w1) * 0x1p-52;
      } else {
        eta[c] = __hiloint2double(v[0][c], v[1][c]);
      }
    }
    logit_terms4(prod, sum, eta, (it & 1) << 31, tab, d);
    if (MODE >= 2) {
      uint32_t lo[4], hi[4];
#pragma unroll
      for (int c = 0; c < 4; c++) digits6(d[c], lo[c], hi[c]);
      const uint32_t l01a = __byte_perm(lo[0], lo[1], 0x5140), l01b = __byte_perm(lo[0], lo[1], 0x7362);
      const uint32_t l23a = __byte_perm(lo[2], lo[3], 0x5140), l23b = __byte_perm(lo[2], lo[3], 0x7362);
      const uint32_t h01 = __byte_perm(hi[0], hi[1], 0x5140), h23 = __byte_perm(hi[2], hi[3], 0x5140);
      acc ^= __byte_perm(l01a, l23a, 0x5410) ^ __byte_perm(l01a, l23a, 0x7632) ^ __byte_perm(l01b, l23b, 0x5410) ^ __byte_perm(l01b, l23b, 0x7632) ^
             __byte_perm(h01, h23, 0x5410) ^ __byte_perm(h01, h23, 0x7632);
    }
#pragma unroll
    for (int c = 0; c < 4; c++) {   // the next iteration's inputs depend on this one's outputs (no hoisting), cheaply
      v[5][c] += __double2loint(d[c]) & 3;
      if (MODE == 0) v[1][c] += __double2loint(d[c]) & 3;
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = prod[0] + prod[1] + prod[2] + prod[3] + sum[0] + sum[1] + sum[2] + sum[3];
  outw[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}


// (3) the same FP64 work while one thread of the CTA keeps the tensor core busy with tcgen05.mma kind::i8 (M = 128, N = 192, K = 32,
// operands anywhere in shared memory): do the scalar FP64 pipe and the 5th-generation tensor core run side by side at full rate?
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int MMA_ON>
__global__ void __launch_bounds__(544, 1) link_beside_mma_kernel(const int* diag, double* out, uint32_t* outw, int iters, long long* cyc, long long* mma_count) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ double tab[32];
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int stop;
  __shared__ __align__(8) unsigned long long bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid < 32) tab[tid] = kExp2Tab32[tid];
  if (tid == 0) { stop = 0; asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  for (uint32_t i = tid * 16; i < 64 * 1024; i += 544 * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0x01010101u, 0x02020202u, 0x03030303u, 0x04040404u);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_base_s;
  if (warp == 0) {
    if (tid == 0 && MMA_ON) {
      const uint32_t base = smem_u32(smem);
      const uint64_t ad = (uint64_t)((base >> 4) & 0x3fff) | ((uint64_t)8 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46);
      const uint64_t bd = (uint64_t)(((base + 32768) >> 4) & 0x3fff) | ((uint64_t)8 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46);
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(192 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      long long n = 0;
      while (!stop) {
#pragma unroll
        for (int u = 0; u < 8; u++)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tbase), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
        n += 8;
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      uint32_t done = 0;
      while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
      mma_count[blockIdx.x] = n;
    }
  } else {
    double prod[4] = {1, 1, 1, 1}, sum[4] = {0, 0, 0, 0};
    int v[6][4];
    const int t = tid - 32;
    for (int k = 0; k < 6; k++)
      for (int c = 0; c < 4; c++) v[k][c] = diag[(t * 6 + k) * 4 + c];
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
      double eta[4], d[4];
#pragma unroll
      for (int c = 0; c < 4; c++) eta[c] = __hiloint2double(v[0][c], v[1][c]);
      logit_terms4(prod, sum, eta, (it & 1) << 31, tab, d);
#pragma unroll
      for (int c = 0; c < 4; c++) v[1][c] += __double2loint(d[c]) & 3;
    }
    const long long t1 = clock64();
    out[blockIdx.x * 512 + t] = prod[0] + prod[1] + prod[2] + prod[3] + sum[0] + sum[1] + sum[2] + sum[3];
    if (t == 0) cyc[blockIdx.x] = t1 - t0;
    asm volatile("bar.sync 1, 512;" ::: "memory");
    if (t == 0) stop = 1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
}

template <typename F>
static double run(F launch, int blocks) {
  long long* dc; cudaMalloc(&dc, blocks * 8);
  launch(dc); cudaDeviceSynchronize();
  launch(dc); cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return -1; }
  long long* h = new long long[blocks]; cudaMemcpy(h, dc, blocks * 8, cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < blocks; i++) mx = h[i] > mx ? h[i] : mx;
  delete[] h; cudaFree(dc);
  return (double)mx;
}

int main() {
  const int blocks = 148;
  double* out; uint32_t* outw; int* diag;
  cudaMalloc(&out, blocks * 1024 * 8); cudaMalloc(&outw, blocks * 1024 * 4); cudaMalloc(&diag, 1024 * 24 * 4);
  int* hd = new int[1024 * 24];
  for (int i = 0; i < 1024 * 24; i++) hd[i] = (i * 2654435761u) % 200001 - 100000;
  cudaMemcpy(diag, hd, 1024 * 24 * 4, cudaMemcpyHostToDevice);
  printf("== DFMA: cycles per warp-level DFMA per SM sub-partition (2.0 = 16 lanes / clk) ==\n");
  for (int warps : {4, 8, 16, 32}) {
    const int iters = 2000;
    double c1 = run([&](long long* dc) { dfma_kernel<1><<<blocks, warps * 32>>>(out, iters, 0.999, 1e-9, dc); }, blocks);
    double c2 = run([&](long long* dc) { dfma_kernel<2><<<blocks, warps * 32>>>(out, iters, 0.999, 1e-9, dc); }, blocks);
    double c4 = run([&](long long* dc) { dfma_kernel<4><<<blocks, warps * 32>>>(out, iters, 0.999, 1e-9, dc); }, blocks);
    double c8 = run([&](long long* dc) { dfma_kernel<8><<<blocks, warps * 32>>>(out, iters, 0.999, 1e-9, dc); }, blocks);
    const double per = warps / 4.0 * iters * 8;   // DFMAs per sub-partition per chain of the ILP set
    printf("%2d warps / SM: ILP1 %.2f (latency %.1f at 1 warp / SMSP)  ILP2 %.2f  ILP4 %.2f  ILP8 %.2f\n", warps, c1 / per, c1 / (iters * 8.0), c2 / (per * 2), c4 / (per * 4), c8 / (per * 8));
  }
  printf("== per-element code of the INT8 kernel's epilogue on register data: cycles per (warp, element) per SM sub-partition ==\n");
  for (int warps : {4, 8, 16, 32}) {
    const int iters = 2000;
    double m0 = run([&](long long* dc) { link_kernel<0><<<blocks, warps * 32>>>(diag, out, outw, iters, dc); }, blocks);
    double m1 = run([&](long long* dc) { link_kernel<1><<<blocks, warps * 32>>>(diag, out, outw, iters, dc); }, blocks);
    double m2 = run([&](long long* dc) { link_kernel<2><<<blocks, warps * 32>>>(diag, out, outw, iters, dc); }, blocks);
    const double per = warps / 4.0 * iters * 4;
    printf("%2d warps / SM: link %.1f   + recombination %.1f   + digits %.1f   (x 1024 elements per sub-partition and tile = %.0f cycles per tile)\n", warps, m0 / per,
           m1 / per, m2 / per, m2 / per * 1024 / 1.0);
  }
  printf("== the link (16 warps / SM) beside a saturated tcgen05.mma kind::i8 stream issued by a 17th warp ==\n");
  {
    long long* mc; cudaMalloc(&mc, blocks * 8); cudaMemset(mc, 0, blocks * 8);
    const int iters = 2000;
    cudaFuncSetAttribute(link_beside_mma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(link_beside_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    double off = run([&](long long* dc) { link_beside_mma_kernel<0><<<blocks, 544, 64 * 1024>>>(diag, out, outw, iters, dc, mc); }, blocks);
    double on = run([&](long long* dc) { link_beside_mma_kernel<1><<<blocks, 544, 64 * 1024>>>(diag, out, outw, iters, dc, mc); }, blocks);
    long long hm[148]; cudaMemcpy(hm, mc, blocks * 8, cudaMemcpyDeviceToHost);
    const double per = 4.0 * iters * 4;
    printf("tensor core idle: %.1f cycles per (warp, element) per sub-partition;  tensor core busy: %.1f  (%.0f MMAs of 128 x 192 x 32 in %.0f cycles = %.0f MAC/clk/SM)\n",
           off / per, on / per, (double)hm[0], on, (double)hm[0] * 128 * 192 * 32 / on);
  }
  return 0;
}
