// FP64 tensor-core (DMMA) variant of the fused cross-chain GLM gradient (see glm.cu for the algorithm and the
// reference citations).  Both contractions of a row tile run on the tensor cores with mma.sync.m8n8k4.f64
// (tcgen05 has no FP64 kind; DMMA is the FP64 tensor path on sm_100a):
//
//   phase A   Eta[TR x TC]  = Xs[TR x Dk] . Bs[Dk x TC]        (A = X tile, row-major;  B = q of the chain tile)
//   link      lp terms, Dm = dl/deta in registers -> Ds[TR x TC] in shared memory
//   phase B   G[Dk x TC]   += Xs^T[Dk x TR] . Ds[TR x TC]      (A = X tile read transposed from the same smem)
//
// The X tile is loaded ONCE per (row tile, chain tile) and serves both products; Eta / Dm never leave the SM.
// Shared-memory leading dimensions are chosen == 4 (mod 16) doubles so that every fragment load (8 rows x 4 k) hits
// 16 distinct 8-byte banks per half-warp.  The next row tile is prefetched with cp.async while the current one is in
// the tensor pipe.
//
// Fragment layout of mma.m8n8k4.f64 (PTX ISA): A a0 = A[lane>>2][lane&3]; B b0 = B[lane&3][lane>>2];
// C c0,c1 = C[lane>>2][2*(lane&3) + {0,1}].
#include "kernels.h"

namespace b200 {

namespace {

constexpr int TR = 64, TC = 64, NT = 256;
constexpr int BS_LD = TC + 4;  // 68 == 4 mod 16
constexpr int DS_LD = TC + 4;

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::); }

__host__ __device__ inline int xs_ld_for(int Dk) {
  int ld = Dk + 4;
  while ((ld & 15) != 4) ld++;
  return ld;
}

template <int LINK>
__device__ __forceinline__ void link_eval(double eta, int yv, double& lpi, double& d) {
  if (LINK == 0) {
    const double sgn = 2.0 * yv - 1.0;
    const double t = sgn * eta;
    const double e = exp(-t);
    if (t > 20.0) {
      lpi = -e;
      d = -e;
    } else if (t < -20.0) {
      lpi = t;
      d = sgn;
    } else {
      lpi = -log1p(e);
      d = sgn * e / (e + 1.0);
    }
  } else {
    const double ee = exp(eta);
    lpi = yv * eta - ee;
    d = yv - ee;
  }
}

// MB: 8-row blocks of G per warp half (Dk <= 16*MB)
template <int LINK, int MB>
__global__ void __launch_bounds__(NT, 1)
glm_dmma_kernel(const double* __restrict__ X, int ldx, const int* __restrict__ y, int N, int Dk, int Dp,
                const double* __restrict__ Q, int ldq, int D, const int* __restrict__ active, int C, int rows_per_group,
                double* __restrict__ Gp, double* __restrict__ LPp) {
  extern __shared__ double smem[];
  const int xs_ld = xs_ld_for(Dk);
  double* Bs = smem;                               // [Dk][BS_LD]
  double* Xs0 = Bs + (size_t)Dk * BS_LD;           // [2][TR][xs_ld]  double-buffered X tile
  double* Ds = Xs0 + (size_t)2 * TR * xs_ld;       // [TR][DS_LD]
  int* ys0 = (int*)(Ds + (size_t)TR * DS_LD);      // [2][TR]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int chain0 = blockIdx.x * TC;
  const int p = blockIdx.y;

  if (active) {
    int any = 0;
    for (int c = tid; c < TC; c += NT)
      if (chain0 + c < C && active[chain0 + c]) any = 1;
    any = __syncthreads_or(any);
    if (!any) return;
  }
  for (int e = tid; e < Dk * TC; e += NT) {
    const int c = e / Dk, k = e - c * Dk;
    double v = 0.0;
    if (chain0 + c < C && k < D) v = Q[(size_t)(chain0 + c) * ldq + k];
    Bs[k * BS_LD + c] = v;
  }
  const long long row_begin = (long long)p * rows_per_group;
  long long row_end = row_begin + rows_per_group;
  if (row_end > N) row_end = N;

  auto prefetch = [&](long long row0, int buf) {
    double* Xs = Xs0 + (size_t)buf * TR * xs_ld;
    int* ys = ys0 + buf * TR;
    for (int e = tid; e < TR * Dk; e += NT) {
      const int r = e / Dk, k = e - r * Dk;
      const long long row = row0 + r;
      if (row < row_end && k < ldx) cp_async8(&Xs[r * xs_ld + k], &X[(size_t)row * ldx + k]);
      else Xs[r * xs_ld + k] = 0.0;
    }
    if (tid < TR) {
      const long long row = row0 + tid;
      ys[tid] = row < row_end ? y[row] : 0;
    }
    cp_async_commit();
  };

  // phase-A tiling: warp -> 32 rows x 16 chains (4 x 2 mma tiles); phase-B tiling: warp -> MB m-blocks x 16 chains
  const int a_rb = (warp & 1) * 32, a_cb = (warp >> 1) * 16;
  const int b_mh = (warp & 1) * MB, b_cb = (warp >> 1) * 16;
  double gacc[MB][2][2];
#pragma unroll
  for (int m = 0; m < MB; m++)
#pragma unroll
    for (int n = 0; n < 2; n++) gacc[m][n][0] = gacc[m][n][1] = 0.0;
  double lpacc[2][2] = {{0, 0}, {0, 0}};

  int buf = 0;
  if (row_begin < row_end) prefetch(row_begin, 0);
  for (long long row0 = row_begin; row0 < row_end; row0 += TR) {
    cp_async_wait_all();
    __syncthreads();  // tile `buf` landed; every warp is done with the previous tile (Ds and the other buffer are free)
    if (row0 + TR < row_end) prefetch(row0 + TR, buf ^ 1);
    const double* Xs = Xs0 + (size_t)buf * TR * xs_ld;
    const int* ys = ys0 + buf * TR;
    // ---- phase A
    double acc[4][2][2];
#pragma unroll
    for (int mi = 0; mi < 4; mi++)
#pragma unroll
      for (int ni = 0; ni < 2; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    for (int k0 = 0; k0 < Dk; k0 += 4) {
      double a[4], b[2];
#pragma unroll
      for (int mi = 0; mi < 4; mi++) a[mi] = Xs[(a_rb + mi * 8 + gid) * xs_ld + k0 + tig];
#pragma unroll
      for (int ni = 0; ni < 2; ni++) b[ni] = Bs[(k0 + tig) * BS_LD + a_cb + ni * 8 + gid];
#pragma unroll
      for (int mi = 0; mi < 4; mi++)
#pragma unroll
        for (int ni = 0; ni < 2; ni++) dmma(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
    }
    // ---- link on the accumulator fragments: element (row = a_rb + mi*8 + gid, chain = a_cb + ni*8 + 2*tig + j)
#pragma unroll
    for (int mi = 0; mi < 4; mi++) {
      const int r = a_rb + mi * 8 + gid;
      const bool valid = row0 + r < row_end;
      const int yv = ys[r];
#pragma unroll
      for (int ni = 0; ni < 2; ni++) {
        double dv[2];
#pragma unroll
        for (int j = 0; j < 2; j++) {
          double lpi, d;
          link_eval<LINK>(acc[mi][ni][j], yv, lpi, d);
          if (!valid) {
            lpi = 0.0;
            d = 0.0;
          }
          lpacc[ni][j] += lpi;
          dv[j] = d;
        }
        *reinterpret_cast<double2*>(&Ds[r * DS_LD + a_cb + ni * 8 + 2 * tig]) = make_double2(dv[0], dv[1]);
      }
    }
    __syncthreads();
    // ---- phase B: G[m][chain] += sum_r Xs[r][m] * Ds[r][chain]
    for (int r0 = 0; r0 < TR; r0 += 4) {
      double b[2];
#pragma unroll
      for (int ni = 0; ni < 2; ni++) b[ni] = Ds[(r0 + tig) * DS_LD + b_cb + ni * 8 + gid];
#pragma unroll
      for (int m = 0; m < MB; m++) {
        const int mcol = (b_mh + m) * 8 + gid;
        const double a = mcol < Dk ? Xs[(r0 + tig) * xs_ld + mcol] : 0.0;
#pragma unroll
        for (int ni = 0; ni < 2; ni++) dmma(gacc[m][ni][0], gacc[m][ni][1], a, b[ni]);
      }
    }
    buf ^= 1;
  }
  // ---- partial gradient of this row group
#pragma unroll
  for (int m = 0; m < MB; m++) {
    const int k = (b_mh + m) * 8 + gid;
    if (k < Dp) {
#pragma unroll
      for (int ni = 0; ni < 2; ni++)
#pragma unroll
        for (int j = 0; j < 2; j++) {
          const int chain = chain0 + b_cb + ni * 8 + 2 * tig + j;
          if (chain < C) Gp[((size_t)p * C + chain) * Dp + k] = gacc[m][ni][j];
        }
    }
  }
  // ---- partial log-likelihood: sum over the 8 row-lanes (gid) by shuffles, then over the two row halves via smem
#pragma unroll
  for (int ni = 0; ni < 2; ni++)
#pragma unroll
    for (int j = 0; j < 2; j++) {
      double v = lpacc[ni][j];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      lpacc[ni][j] = v;
    }
  __syncthreads();
  double* red = Ds;  // [2][TC]
  if (gid == 0) {
#pragma unroll
    for (int ni = 0; ni < 2; ni++)
#pragma unroll
      for (int j = 0; j < 2; j++) red[(warp & 1) * TC + a_cb + ni * 8 + 2 * tig + j] = lpacc[ni][j];
  }
  __syncthreads();
  if (tid < TC && chain0 + tid < C) LPp[(size_t)p * C + chain0 + tid] = red[tid] + red[TC + tid];
}

}  // namespace

int launch_glm_dmma(const ModelData& m, const GradBatch& b, const GlmWorkspace& ws, cudaStream_t stream) {
  const int Dp = ws.Dp, C = b.C, P = ws.P;
  const int Dk = (m.D + 7) & ~7;
  int rows_per_group = (m.N + P - 1) / P;
  rows_per_group = (rows_per_group + TR - 1) / TR * TR;
  const int xs_ld = xs_ld_for(Dk);
  const size_t smem = ((size_t)Dk * BS_LD + (size_t)2 * TR * xs_ld + (size_t)TR * DS_LD) * sizeof(double) + 2 * TR * sizeof(int);
  dim3 grid((C + TC - 1) / TC, P);
#define LAUNCH(LINK, MB)                                                                                              \
  {                                                                                                                   \
    auto k = glm_dmma_kernel<LINK, MB>;                                                                               \
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                                  \
    k<<<grid, NT, smem, stream>>>(m.X, m.ldx, m.yi, m.N, Dk, Dp, b.Q, b.ld, m.D, b.active, C, rows_per_group, ws.Gp, ws.LPp); \
  }
  const bool logistic = m.kind == MODEL_LOGISTIC;
  if (Dk <= 64) {
    if (logistic) LAUNCH(0, 4) else LAUNCH(1, 4)
  } else {  // Dk <= 112 (larger designs use the DFMA kernel of glm.cu: the double-buffered tile would not fit in smem)
    if (logistic) LAUNCH(0, 7) else LAUNCH(1, 7)
  }
#undef LAUNCH
  return 1;
}

}  // namespace b200
