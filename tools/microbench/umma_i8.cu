// Layout check + issue-rate microbenchmark of tcgen05.mma.kind::i8 (INT8 x INT8 -> INT32 in tensor memory) on sm_100a, for the
// Ozaki-split GLM gradient kernel (csrc/glm_i8.cu).  What it settles on hardware before the kernel relies on it:
//  (1) the shared-memory matrix descriptors without swizzle ("interleaved" core matrices of 8 x 16 bytes):
//        K-major  operand: core matrix = 8 MN-rows x 16 K-bytes, LBO = stride of the next 16 K-bytes, SBO = stride of the next 8 rows
//        MN-major operand: core matrix = 8 K-rows  x 16 MN-bytes, SBO = stride of the next 16 MN-bytes, LBO = stride of the next 8 K-rows
//      (the same X tile [row / 8][col / 16][8][16] then serves X . q as a K-major A and X^T . dm as an MN-major A);
//  (2) instruction descriptor bits for kind::i8 (signed x signed, S32 accumulators, a_major / b_major, M = 128, N <= 256);
//  (3) accumulation of several products into overlapping column windows of one accumulator (the slice diagonals);
//  (4) tcgen05.ld.32x32b: thread = TMEM lane (row of D), registers = consecutive columns;
//  (5) the tensor time of one row tile of the kernel (15 slice products per contraction, N = 160 .. 32) with nothing else running.
// Each layout test is run with the assumed (LBO, SBO) and with the two swapped, and reports which one reproduces the host product.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_i8 umma_i8.cu     Run: ./umma_i8
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

struct MmaOp {
  uint32_t a_off, b_off;      // byte offsets into the A / B regions of shared memory (first K step)
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo, a_kstep, b_kstep;   // bytes
  uint32_t idesc, d_col, accumulate, ksteps;
};
constexpr int kMaxOps = 64;
struct Args {
  const uint8_t* a_img; const uint8_t* b_img; uint32_t a_bytes, b_bytes;
  int32_t* D; int n_cols; int n_ops; int repeat; long long* cycles; int* err;
  MmaOp ops[kMaxOps];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version of sm_100
  return d;                 // base offset 0, LBO mode 0, layout type 0 = no swizzle
}
__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) umma_kernel(const Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) unsigned long long bar;
  uint8_t* As = smem;
  uint8_t* Bs = smem + ((a.a_bytes + 1023) & ~1023u);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (uint32_t i = tid * 16; i < a.a_bytes; i += 128 * 16) *reinterpret_cast<uint4*>(As + i) = *reinterpret_cast<const uint4*>(a.a_img + i);
  for (uint32_t i = tid * 16; i < a.b_bytes; i += 128 * 16) *reinterpret_cast<uint4*>(Bs + i) = *reinterpret_cast<const uint4*>(a.b_img + i);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes of the operands -> visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_base_s;
  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    t0 = clock64();
    for (int rep = 0; rep < a.repeat; rep++) {
      for (int o = 0; o < a.n_ops; o++) {
        const MmaOp& op = a.ops[o];
        for (uint32_t k = 0; k < op.ksteps; k++) {
          const uint64_t ad = make_desc(smem_u32(As) + op.a_off + k * op.a_kstep, op.a_lbo, op.a_sbo);
          const uint64_t bd = make_desc(smem_u32(Bs) + op.b_off + k * op.b_kstep, op.b_lbo, op.b_sbo);
          mma_i8(tbase + op.d_col, ad, bd, op.idesc, (rep > 0 || k > 0 || op.accumulate) ? 1u : 0u);
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // everybody waits for the tensor core (bounded: a wrong descriptor must not hang the box)
  {
    uint32_t done = 0;
    for (long long spin = 0; spin < (1ll << 24) && !done; spin++) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    }
    if (!done && tid == 0) *a.err = 1;
  }
  if (tid == 0) {
    t1 = clock64();
    if (a.cycles) a.cycles[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (a.D && blockIdx.x == 0) {
    for (int c0 = 0; c0 < a.n_cols; c0 += 8) {
      uint32_t v[8];
      const uint32_t taddr = tbase + ((uint32_t)(warp * 32) << 16) + c0;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; j++) a.D[(size_t)tid * a.n_cols + c0 + j] = (int32_t)v[j];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
}


// issue-rate kernel: the MMA stream of one row tile with every descriptor a compile-time offset from two bases (what the real kernel
// issues): NSL slice products, concatenated along N, 4 K steps each; MODE 0 = first product (K-major), 1 = second (MN-major)
__host__ __device__ constexpr uint32_t idesc_c(int M, int N, int a_mn, int b_mn) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
template <int MODE, int NSL, int CNT>
__global__ void __launch_bounds__(128, 1) rate_kernel(int repeat, long long* cycles, int* err) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) unsigned long long bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (uint32_t i = tid * 16; i < 200 * 1024; i += 128 * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0x01010101u, 0x02020202u, 0x03030303u, 0x04040404u);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_base_s;
  if (tid == 0) {
    const uint32_t abase = smem_u32(smem), bbase = abase + 96 * 1024;
    const uint64_t ad0 = MODE == 0 ? make_desc(abase, 128, 1024) : make_desc(abase, 1024, 128);
    const uint64_t bd0 = MODE == 0 ? make_desc(bbase, 128, 1024) : make_desc(bbase, 128, 2048);
    const long long t0 = clock64();
    for (int rep = 0; rep < repeat; rep++) {
#pragma unroll
      for (int i = 0; i < NSL; i++) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const uint32_t aoff = i * 16384 + k * (MODE == 0 ? 256 : 4096), boff = k * (MODE == 0 ? 256 : 512);
          mma_i8(tbase + i * CNT, ad0 + (aoff >> 4), bd0 + (boff >> 4), idesc_c(128, (NSL - i) * CNT, MODE, MODE), 1u);
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t done = 0;
    for (long long spin = 0; spin < (1ll << 26) && !done; spin++)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    if (!done) *err = 1;
    cycles[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
}
template <int MODE, int NSL, int CNT>
static void run_rate(const char* name) {
  int* d_err; CK(cudaMalloc(&d_err, 4)); CK(cudaMemset(d_err, 0, 4));
  long long* dC; CK(cudaMalloc(&dC, 148 * 8));
  auto k = rate_kernel<MODE, NSL, CNT>;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int repeat = 400;
  for (int it = 0; it < 2; it++) { k<<<148, 128, 200 * 1024>>>(repeat, dC, d_err); CK(cudaDeviceSynchronize()); }
  std::vector<long long> cyc(148); int err = 0;
  CK(cudaMemcpy(cyc.data(), dC, 148 * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost));
  long long mx = 0; for (auto c : cyc) mx = c > mx ? c : mx;
  const double macs = 0.5 * NSL * (NSL + 1) * 128.0 * CNT * 128;
  printf("unrolled %s, %d slices, %d chains: %.0f cycles per row tile, %.0f MAC/clk/SM, timeout=%d\n", name, NSL, CNT, (double)mx / repeat, macs * repeat / (double)mx, err);
  CK(cudaFree(dC)); CK(cudaFree(d_err));
}

static uint32_t idesc_i8(int M, int N, int a_mn, int b_mn) {
  uint32_t d = 0;
  d |= 2u << 4;                 // accumulator format S32
  d |= 1u << 7;                 // A: signed 8 bit
  d |= 1u << 10;                // B: signed 8 bit
  d |= (uint32_t)a_mn << 15;    // A major: 0 = K, 1 = MN
  d |= (uint32_t)b_mn << 16;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}

constexpr int NS = 5;          // slices
constexpr int TRW = 128;       // rows of a row tile
constexpr int DK = 128;        // padded design columns
constexpr int CN = 32;         // chains of a chain tile

// X tile image [slice][row / 8][col / 16][8][16]
static size_t x_idx(int s, int r, int c) { return ((((size_t)s * (TRW / 8) + r / 8) * (DK / 16) + c / 16) * 8 + r % 8) * 16 + c % 16; }
// q image (K-major B of the first product): [slice][chain / 8][col / 16][8][16]
static size_t q_idx(int s, int ch, int c) { return ((((size_t)s * (CN / 8) + ch / 8) * (DK / 16) + c / 16) * 8 + ch % 8) * 16 + c % 16; }
// dm image (MN-major B of the second product): [slice][chain / 16][row / 8][8][16]
static size_t dm_idx(int s, int r, int ch) { return ((((size_t)s * (CN / 16) + ch / 16) * (TRW / 8) + r / 8) * 8 + r % 8) * 16 + ch % 16; }

static int run(Args& a, size_t smem, int grid, std::vector<int32_t>* out, std::vector<long long>* cyc, float* ms) {
  int* d_err; CK(cudaMalloc(&d_err, 4)); CK(cudaMemset(d_err, 0, 4));
  int32_t* dD = nullptr; long long* dC = nullptr;
  if (out) { CK(cudaMalloc(&dD, (size_t)128 * a.n_cols * 4)); CK(cudaMemset(dD, 0xff, (size_t)128 * a.n_cols * 4)); }
  CK(cudaMalloc(&dC, grid * 8));
  a.D = dD; a.cycles = dC; a.err = d_err;
  CK(cudaFuncSetAttribute(umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  umma_kernel<<<grid, 128, smem>>>(a);
  CK(cudaEventRecord(e1));
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("  kernel failed: %s\n", cudaGetErrorString(e)); exit(3); }
  if (ms) CK(cudaEventElapsedTime(ms, e0, e1));
  int err = 0; CK(cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost));
  if (out) { out->resize((size_t)128 * a.n_cols); CK(cudaMemcpy(out->data(), dD, out->size() * 4, cudaMemcpyDeviceToHost)); CK(cudaFree(dD)); }
  if (cyc) { cyc->resize(grid); CK(cudaMemcpy(cyc->data(), dC, grid * 8, cudaMemcpyDeviceToHost)); }
  CK(cudaFree(dC)); CK(cudaFree(d_err));
  return err;
}

int main() {
  srand(12345);
  std::vector<int8_t> X((size_t)NS * TRW * DK), Q((size_t)NS * CN * DK), DM((size_t)NS * TRW * CN);
  for (auto& v : X) v = (int8_t)(rand() % 255 - 127);
  for (auto& v : Q) v = (int8_t)(rand() % 255 - 127);
  for (auto& v : DM) v = (int8_t)(rand() % 255 - 127);
  auto x = [&](int s, int r, int c) { return (int)X[((size_t)s * TRW + r) * DK + c]; };
  auto q = [&](int s, int ch, int c) { return (int)Q[((size_t)s * CN + ch) * DK + c]; };
  auto dm = [&](int s, int r, int ch) { return (int)DM[((size_t)s * TRW + r) * CN + ch]; };
  std::vector<uint8_t> ximg((size_t)NS * TRW * DK), qimg((size_t)NS * CN * DK), dimg((size_t)NS * TRW * CN);
  for (int s = 0; s < NS; s++)
    for (int r = 0; r < TRW; r++)
      for (int c = 0; c < DK; c++) ximg[x_idx(s, r, c)] = (uint8_t)X[((size_t)s * TRW + r) * DK + c];
  for (int s = 0; s < NS; s++)
    for (int ch = 0; ch < CN; ch++)
      for (int c = 0; c < DK; c++) qimg[q_idx(s, ch, c)] = (uint8_t)Q[((size_t)s * CN + ch) * DK + c];
  for (int s = 0; s < NS; s++)
    for (int r = 0; r < TRW; r++)
      for (int ch = 0; ch < CN; ch++) dimg[dm_idx(s, r, ch)] = (uint8_t)DM[((size_t)s * TRW + r) * CN + ch];
  uint8_t *dX, *dQ, *dDm;
  CK(cudaMalloc(&dX, ximg.size())); CK(cudaMalloc(&dQ, qimg.size())); CK(cudaMalloc(&dDm, dimg.size()));
  CK(cudaMemcpy(dX, ximg.data(), ximg.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dQ, qimg.data(), qimg.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dDm, dimg.data(), dimg.size(), cudaMemcpyHostToDevice));
  const size_t smem = ((ximg.size() + 1023) & ~(size_t)1023) + qimg.size() + 1024;

  // expected results: diagonal accumulators, column = diag * CN + chain
  std::vector<long long> expA((size_t)128 * NS * CN, 0), expB((size_t)128 * NS * CN, 0);
  for (int i = 0; i < NS; i++)
    for (int j = 0; i + j < NS; j++)
      for (int r = 0; r < TRW; r++)
        for (int ch = 0; ch < CN; ch++) {
          long long s = 0;
          for (int c = 0; c < DK; c++) s += x(i, r, c) * q(j, ch, c);
          expA[(size_t)r * NS * CN + (i + j) * CN + ch] += s;
        }
  for (int i = 0; i < NS; i++)
    for (int j = 0; i + j < NS; j++)
      for (int c = 0; c < DK; c++)
        for (int ch = 0; ch < CN; ch++) {
          long long s = 0;
          for (int r = 0; r < TRW; r++) s += x(i, r, c) * dm(j, r, ch);
          expB[(size_t)c * NS * CN + (i + j) * CN + ch] += s;
        }

  auto build_A = [&](Args& a, bool swap_a, bool swap_b) {   // Eta diagonals: X_i (K-major A) x [q_0 .. q_{NS-1-i}] (K-major B)
    a.n_ops = 0;
    for (int i = 0; i < NS; i++) {
      MmaOp& op = a.ops[a.n_ops++];
      op.a_off = (uint32_t)x_idx(i, 0, 0); op.b_off = 0;
      op.a_lbo = 128; op.a_sbo = (DK / 16) * 128; op.a_kstep = 256;
      op.b_lbo = 128; op.b_sbo = (DK / 16) * 128; op.b_kstep = 256;
      if (swap_a) std::swap(op.a_lbo, op.a_sbo);
      if (swap_b) std::swap(op.b_lbo, op.b_sbo);
      op.idesc = idesc_i8(128, (NS - i) * CN, 0, 0);
      op.d_col = i * CN; op.accumulate = i > 0; op.ksteps = DK / 32;
    }
  };
  auto build_B = [&](Args& a, bool swap_a, bool swap_b) {   // G diagonals: X_i^T (MN-major A) x [dm_0 .. dm_{NS-1-i}] (MN-major B)
    a.n_ops = 0;
    for (int i = 0; i < NS; i++) {
      MmaOp& op = a.ops[a.n_ops++];
      op.a_off = (uint32_t)x_idx(i, 0, 0); op.b_off = 0;
      op.a_sbo = 128; op.a_lbo = (DK / 16) * 128; op.a_kstep = 4 * (DK / 16) * 128;
      op.b_sbo = (TRW / 8) * 128; op.b_lbo = 128; op.b_kstep = 4 * 128;
      if (swap_a) std::swap(op.a_lbo, op.a_sbo);
      if (swap_b) std::swap(op.b_lbo, op.b_sbo);
      op.idesc = idesc_i8(128, (NS - i) * CN, 1, 1);
      op.d_col = i * CN; op.accumulate = i > 0; op.ksteps = TRW / 32;
    }
  };
  auto check = [&](const char* name, const std::vector<int32_t>& got, const std::vector<long long>& exp) {
    size_t bad = 0, first = (size_t)-1;
    for (size_t i = 0; i < exp.size(); i++)
      if ((long long)got[i] != exp[i]) { if (!bad) first = i; bad++; }
    printf("  %s: %zu / %zu mismatches", name, bad, exp.size());
    if (bad) printf("  (first at row %zu col %zu: got %d expected %lld)", first / (NS * CN), first % (NS * CN), got[first], exp[first]);
    printf("\n");
    return bad == 0;
  };

  printf("== layout tests (M = 128, K = 128, N = %d .. %d, %d slices, diagonal accumulation) ==\n", NS * CN, CN, NS);
  for (int sa = 0; sa < 2; sa++)
    for (int sb = 0; sb < 2; sb++) {
      Args a; memset(&a, 0, sizeof a);
      a.a_img = dX; a.a_bytes = (uint32_t)ximg.size(); a.b_img = dQ; a.b_bytes = (uint32_t)qimg.size();
      a.n_cols = NS * CN; a.repeat = 1;
      build_A(a, sa, sb);
      std::vector<int32_t> out;
      const int err = run(a, smem, 1, &out, nullptr, nullptr);
      printf("first product (K-major A, K-major B), A %s, B %s, timeout=%d\n", sa ? "LBO/SBO swapped" : "as assumed", sb ? "LBO/SBO swapped" : "as assumed", err);
      check("Eta diagonals", out, expA);
    }
  for (int sa = 0; sa < 2; sa++)
    for (int sb = 0; sb < 2; sb++) {
      Args a; memset(&a, 0, sizeof a);
      a.a_img = dX; a.a_bytes = (uint32_t)ximg.size(); a.b_img = dDm; a.b_bytes = (uint32_t)dimg.size();
      a.n_cols = NS * CN; a.repeat = 1;
      build_B(a, sa, sb);
      std::vector<int32_t> out;
      const int err = run(a, smem, 1, &out, nullptr, nullptr);
      printf("second product (MN-major A, MN-major B), A %s, B %s, timeout=%d\n", sa ? "LBO/SBO swapped" : "as assumed", sb ? "LBO/SBO swapped" : "as assumed", err);
      check("G diagonals", out, expB);
    }

  printf("== tensor time (148 CTAs, one per SM; the MMA stream of R row tiles, nothing else running) ==\n");
  for (int which = 0; which < 3; which++) {
    Args a; memset(&a, 0, sizeof a);
    a.a_img = dX; a.a_bytes = (uint32_t)ximg.size();
    a.n_cols = NS * CN; a.repeat = 200;
    const char* name;
    if (which == 0) { a.b_img = dQ; a.b_bytes = (uint32_t)qimg.size(); build_A(a, false, false); name = "first product  (5 MMAs x 4 K steps, N = 160..32)"; }
    else if (which == 1) { a.b_img = dDm; a.b_bytes = (uint32_t)dimg.size(); build_B(a, false, false); name = "second product (5 MMAs x 4 K steps, N = 160..32)"; }
    else {   // the same MAC count as separate N = 32 products (15 MMAs x 4 K steps): what concatenation along N buys
      a.b_img = dQ; a.b_bytes = (uint32_t)qimg.size();
      a.n_ops = 0;
      for (int i = 0; i < NS; i++)
        for (int j = 0; i + j < NS; j++) {
          MmaOp& op = a.ops[a.n_ops++];
          op.a_off = (uint32_t)x_idx(i, 0, 0); op.b_off = (uint32_t)q_idx(j, 0, 0);
          op.a_lbo = 128; op.a_sbo = (DK / 16) * 128; op.a_kstep = 256;
          op.b_lbo = 128; op.b_sbo = (DK / 16) * 128; op.b_kstep = 256;
          op.idesc = idesc_i8(128, CN, 0, 0);
          op.d_col = (i + j) * CN; op.accumulate = 1; op.ksteps = DK / 32;
        }
      name = "first product as 15 separate N = 32 MMAs x 4 K steps";
    }
    std::vector<long long> cyc; float ms = 0;
    run(a, smem, 148, nullptr, &cyc, &ms);   // warm-up
    const int err = run(a, smem, 148, nullptr, &cyc, &ms);
    long long mx = 0; for (auto c : cyc) mx = c > mx ? c : mx;
    const double macs = 15.0 * 128 * CN * 128;
    printf("%s: %.0f cycles per row tile (max over CTAs), %.1f MAC/clk/SM, kernel %.3f ms, timeout=%d\n", name, (double)mx / a.repeat,
           macs * a.repeat / (double)mx, ms, err);
  }
  run_rate<0, 5, 32>("first product ");
  run_rate<1, 5, 32>("second product");
  run_rate<0, 5, 48>("first product ");
  run_rate<1, 5, 48>("second product");
  run_rate<0, 6, 32>("first product ");
  run_rate<1, 6, 32>("second product");
  run_rate<0, 4, 64>("first product ");
  run_rate<1, 4, 64>("second product");
  return 0;
}
