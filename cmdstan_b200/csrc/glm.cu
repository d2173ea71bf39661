// Cross-chain GLM log-density + gradient for the lock-step engine (configs c3 logistic, c5 poisson).
//
// Reference semantics: stan::math::bernoulli_logit_glm_lpmf (M/prim/prob/bernoulli_logit_glm_lpmf.hpp:51-167, cutoff
// +-20 branches :121-126,:137-142) and poisson_log_glm_lpmf (M/prim/prob/poisson_log_glm_lpmf.hpp:53-163), with the
// priors of SURVEY.md Appendix E (alpha ~ N(0, 2.5 | 1), beta ~ N(0,1)), propto = true.
//
// The reference evaluates ONE chain: eta = X beta (N x K GEMV), link, X^T d (second GEMV) -- two passes over X per
// gradient per chain.  Here all C chains advance in lock-step, so the two GEMVs become two GEMMs stacked across chains
//     Eta[N x C] = Xp[N x Dp] . B[Dp x C] ,   G[Dp x C] = Xp^T . Dm[N x C]
// (Xp = [1 | X | 0-pad], so the intercept is column 0 and sum(d) is row 0 of G), fused per row tile: a CTA loads a
// tile of rows once, forms Eta for its chain tile, applies the link in registers, and immediately contracts the
// derivative tile back against the same X tile.  Eta and Dm never touch HBM; X is streamed once per chain tile.
//
// Determinism: rows are split into P contiguous groups; each CTA reduces its own group in a fixed order and the groups
// are summed in index order by glm_reduce_kernel, so chain c's (lp, grad) depends only on q_c.
//
// Algorithmic work per chain-gradient: 4 N Dp flop + N link evaluations (exp [+ log1p, div]).
// variant 0 (this file; designs with more than 112 columns): register-tiled DFMA.  variant 1 (glm_dmma.cu, default): FP64
// tensor cores (mma.sync m8n8k4 DMMA), persistent over fixed row groups, compacted chain tiles.
#include <cstdio>

#include "kernels.h"

namespace b200 {

constexpr int TR = 64;    // rows per tile
constexpr int TC = 64;    // chains per tile
constexpr int NT = 256;   // threads per CTA
constexpr int kMaxDp = 176;

int glm_max_dim() { return kMaxDp; }

// per-observation log-likelihood term and d(term)/d(eta)
template <int LINK>
__device__ __forceinline__ void link_eval(double eta, int yv, double& lpi, double& d) {
  if (LINK == 0) {  // bernoulli-logit
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
  } else {  // poisson-log
    const double ee = exp(eta);
    lpi = yv * eta - ee;
    d = yv - ee;
  }
}

template <int LINK, int MK>
__global__ void __launch_bounds__(NT, 1)
glm_partial_kernel(const double* __restrict__ X, int ldx, const int* __restrict__ y, int N, int Dp, const double* __restrict__ Q,
                   int ldq, int D, const int* __restrict__ active, int C, long long n_row_tiles, int P, int cap,
                   double* __restrict__ Gp, double* __restrict__ LPp) {
  B200_DYN_SMEM(smem);
  const int xs_ld = Dp + 1;
  double* Bs = smem;                        // [Dp][TC]
  double* Xs = Bs + (size_t)Dp * TC;        // [TR][Dp+1]
  double* Ds = Xs + (size_t)TR * xs_ld;     // [TR][TC]
  int* ys = (int*)(Ds + (size_t)TR * TC);   // [TR]
  const int tid = threadIdx.x;
  const int chain0 = blockIdx.x * TC;
  const int p = blockIdx.y;

  // skip chain tiles without an active chain (tail of a run)
  int any = 0;
  if (active) {
    for (int c = tid; c < TC; c += NT)
      if (chain0 + c < C && active[chain0 + c]) any = 1;
    any = __syncthreads_or(any);
    if (!any) return;
  }
  // B tile: Bs[k][c] = q_c[k]
  for (int e = tid; e < Dp * TC; e += NT) {
    const int c = e / Dp, k = e - c * Dp;  // consecutive threads read consecutive k of one chain (coalesced)
    double v = 0.0;
    if (chain0 + c < C && k < D) v = Q[(size_t)(chain0 + c) * ldq + k];
    Bs[k * TC + c] = v;
  }
  const int ty = tid >> 4, tx = tid & 15;
  const int r0 = ty * 4, c0 = tx * 4;
  double lpacc[4] = {0, 0, 0, 0};
  double gacc[MK][4];
#pragma unroll
  for (int m = 0; m < MK; m++)
#pragma unroll
    for (int c = 0; c < 4; c++) gacc[m][c] = 0.0;

  // fixed row group p: row tiles [p T / P, (p+1) T / P) -- a function of N only (same partition as the DMMA kernel)
  const long long row_begin = ((long long)p * n_row_tiles / P) * TR;
  long long row_end = ((long long)(p + 1) * n_row_tiles / P) * TR;
  if (row_end > N) row_end = N;
  for (long long row0 = row_begin; row0 < row_end; row0 += TR) {
    __syncthreads();  // previous tile fully consumed (and Bs visible on the first pass)
    for (int e = tid; e < TR * Dp; e += NT) {
      const int r = e / Dp, k = e - r * Dp;
      const long long row = row0 + r;
      Xs[r * xs_ld + k] = row < row_end ? X[(size_t)row * ldx + k] : 0.0;
    }
    if (tid < TR) {
      const long long row = row0 + tid;
      ys[tid] = row < row_end ? y[row] : 0;
    }
    __syncthreads();
    // ---- Eta tile: 4 rows x 4 chains per thread
    double acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int c = 0; c < 4; c++) acc[r][c] = 0.0;
    for (int k = 0; k < Dp; k++) {
      const double2 b01 = *reinterpret_cast<const double2*>(&Bs[k * TC + c0]);
      const double2 b23 = *reinterpret_cast<const double2*>(&Bs[k * TC + c0 + 2]);
      const double bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
      for (int r = 0; r < 4; r++) {
        const double xv = Xs[(r0 + r) * xs_ld + k];
#pragma unroll
        for (int c = 0; c < 4; c++) acc[r][c] += xv * bv[c];
      }
    }
    // ---- link: log-likelihood terms and derivative tile
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const bool valid = row0 + r0 + r < row_end;
      const int yv = ys[r0 + r];
      double dv[4];
#pragma unroll
      for (int c = 0; c < 4; c++) {
        double lpi, d;
        link_eval<LINK>(acc[r][c], yv, lpi, d);
        if (!valid) {
          lpi = 0.0;
          d = 0.0;
        }
        lpacc[c] += lpi;
        dv[c] = d;
      }
      *reinterpret_cast<double2*>(&Ds[(r0 + r) * TC + c0]) = make_double2(dv[0], dv[1]);
      *reinterpret_cast<double2*>(&Ds[(r0 + r) * TC + c0 + 2]) = make_double2(dv[2], dv[3]);
    }
    __syncthreads();
    // ---- G tile += X_tile^T . D_tile : thread owns k = ty + 16 m (m < MK) x 4 chains
    for (int r = 0; r < TR; r++) {
      const double2 d01 = *reinterpret_cast<const double2*>(&Ds[r * TC + c0]);
      const double2 d23 = *reinterpret_cast<const double2*>(&Ds[r * TC + c0 + 2]);
      const double dv[4] = {d01.x, d01.y, d23.x, d23.y};
#pragma unroll
      for (int m = 0; m < MK; m++) {
        const int k = ty + 16 * m;
        const double xv = k < Dp ? Xs[r * xs_ld + k] : 0.0;
#pragma unroll
        for (int c = 0; c < 4; c++) gacc[m][c] += xv * dv[c];
      }
    }
  }
  // ---- write the partials of this row group
#pragma unroll
  for (int m = 0; m < MK; m++) {
    const int k = ty + 16 * m;
    if (k < Dp) {
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const int chain = chain0 + c0 + c;
        if (chain < C) Gp[((size_t)p * cap + chain) * Dp + k] = gacc[m][c];
      }
    }
  }
  __syncthreads();
  double* red = Ds;  // [16][TC]
#pragma unroll
  for (int c = 0; c < 4; c++) red[ty * TC + c0 + c] = lpacc[c];
  __syncthreads();
  if (tid < TC && chain0 + tid < C) {
    double s = 0.0;
    for (int t = 0; t < 16; t++) s += red[t * TC + tid];
    LPp[(size_t)p * cap + chain0 + tid] = s;
  }
}

// R[chain][0] = sum_p LPp[p][slot] ; R[chain][1+k] = sum_p Gp[p][slot][k]   (fixed order over row groups).
// slot -> chain through the compacted list of requesting chains (list == nullptr: identity, `active` flags respected).
__global__ void glm_reduce_kernel(const double* __restrict__ Gp, const double* __restrict__ LPp, int P, int cap, int Dp,
                                  const int* __restrict__ list, const int* __restrict__ n_list, int n_fixed,
                                  const int* __restrict__ active, double* __restrict__ R) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int w = 1 + Dp;
  const int n = n_list ? *n_list : n_fixed;
  if (idx >= (long long)n * w) return;
  const int slot = (int)(idx / w), j = (int)(idx - (long long)slot * w);
  const int c = list ? list[slot] : slot;
  if (!list && active && !active[c]) return;
  double s = 0.0;
  if (j == 0)
    for (int p = 0; p < P; p++) s += LPp[(size_t)p * cap + slot];
  else
    for (int p = 0; p < P; p++) s += Gp[((size_t)p * cap + slot) * Dp + (j - 1)];
  R[(size_t)c * w + j] = s;
}

// priors + sign convention: zV = -(loglik + prior), G = -(X^T d + prior gradient).  One warp per chain.
template <int LINK>
__global__ void glm_finish_kernel(const double* __restrict__ R, int Dp, const double* __restrict__ Q, int ldq, int D,
                                  const int* __restrict__ active, int C, double* __restrict__ G, double* __restrict__ zV) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= C) return;
  const int c = warp;
  if (active && !active[c]) return;
  const double* r = R + (size_t)c * (1 + Dp);
  const double* q = Q + (size_t)c * ldq;
  double* g = G + (size_t)c * ldq;
  const double alpha_scale = LINK == 0 ? 2.5 : 1.0;
  double sb = 0.0;
  for (int k = lane; k < D; k += 32) {
    const double qk = q[k];
    double gk = r[1 + k];
    if (k == 0) {
      const double za = qk / alpha_scale;
      gk -= za / alpha_scale;
    } else {
      sb += qk * qk;
      gk += -qk;
    }
    g[k] = -gk;
  }
  sb = warp_sum(sb);
  if (lane == 0) {
    const double za = q[0] / alpha_scale;
    double lp = r[0] + (-0.5 * za * za - 0.5 * sb);
    if (LINK == 1 && !isfinite(r[1])) lp = -INFINITY;  // poisson_log_glm: non-finite sum(d) => domain_error => V = +inf
    zV[c] = -lp;
  }
}

size_t glm_workspace_bytes(int N, int D, int C, int sm_count, GlmWorkspace* ws) {
  const int Dp = (D + 7) & ~7;
  const long long T = ((long long)N + TR - 1) / TR;
  int P = (int)(T < sm_count ? T : sm_count);  // one row group per SM; depends on N only (batch-independent sums)
  if (P < 1) P = 1;
  ws->P = P;
  ws->C = C;
  ws->Dp = Dp;
  return ((size_t)P * C * Dp + (size_t)P * C + (size_t)C * (1 + Dp)) * sizeof(double);
}

void glm_workspace_bind(GlmWorkspace* ws, void* base) {
  ws->Gp = (double*)base;
  ws->LPp = ws->Gp + (size_t)ws->P * ws->C * ws->Dp;
  ws->R = ws->LPp + (size_t)ws->P * ws->C;
}

int launch_glm_dmma(const ModelData& m, const GradBatch& b, const GlmWorkspace& ws, cudaStream_t stream);  // glm_dmma.cu
int glm_tc_max_dk();

template <int LINK>
static int launch_glm_t(const ModelData& m, const GradBatch& b, const GlmWorkspace& ws, allreduce_fn ar, void* ar_user,
                        cudaStream_t stream) {
  const int Dp = ws.Dp, C = b.C;
  int P = ws.P;
  int launches = 0;
#ifdef B200_SIMT_EMUL   // the CPU rig has no stand-in for tcgen05 / tensor memory: the INT8 kernel is validated on hardware only
  const bool i8 = false;
#else
  const bool i8 = LINK == 0 && ws.variant == 7 && ws.i8 != nullptr;
#endif
  const bool tc = i8 || (ws.variant >= 1 && Dp <= glm_tc_max_dk());
  if (i8) {  // INT8 digit planes on tcgen05 (glm_i8.cu): its own, coarser row groups
#ifndef B200_SIMT_EMUL
    launches += launch_glm_i8(m, b, ws, stream);
    P = glm_i8_row_groups(ws.i8);
#endif
  } else if (tc) {
    launches += launch_glm_dmma(m, b, ws, stream);
  } else {
    const long long T = ((long long)m.N + TR - 1) / TR;
    const size_t smem = ((size_t)Dp * TC + (size_t)TR * (Dp + 1) + (size_t)TR * TC) * sizeof(double) + TR * sizeof(int);
    dim3 grid((C + TC - 1) / TC, P);
    if (Dp <= 112) {
      auto k = glm_partial_kernel<LINK, 7>;
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      B200_LAUNCH(k, grid, NT, smem, stream, m.X, m.ldx, m.yi, m.N, Dp, b.Q, b.ld, m.D, b.active, C, T, P, ws.C, ws.Gp, ws.LPp);
    } else {
      auto k = glm_partial_kernel<LINK, 11>;
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      B200_LAUNCH(k, grid, NT, smem, stream, m.X, m.ldx, m.yi, m.N, Dp, b.Q, b.ld, m.D, b.active, C, T, P, ws.C, ws.Gp, ws.LPp);
    }
    launches++;
  }
  // the DMMA kernel works on the compacted list of requesting chains, the DFMA kernel on all chains + active flags
  const int* list = tc ? b.list : nullptr;
  const int* n_list = tc ? b.n_list : nullptr;
  const long long tot = (long long)C * (1 + Dp);
  B200_LAUNCH(glm_reduce_kernel, (unsigned)((tot + 255) / 256), 256, 0, stream, ws.Gp, ws.LPp, P, ws.C, Dp, list, n_list, C, b.active, ws.R);
  launches++;
  if (ar) {  // row sharding: the exchange step of the leapfrog.  A failed collective must stop the run: the finishing kernel
             // would read un-reduced partials and the ranks' control flow would separate
    const int rc = ar(ws.R, (long long)ws.C * (1 + Dp), (void*)stream, ar_user);
    if (rc != 0) return -(rc > 0 ? rc : 1);
  }
  B200_LAUNCH(glm_finish_kernel<LINK>, (C * 32 + 127) / 128, 128, 0, stream, ws.R, Dp, b.Q, b.ld, m.D, b.active, C, b.G, b.zV);
  launches++;
  return launches;
}

int launch_glm_grad(const ModelData& m, const GradBatch& b, const GlmWorkspace& ws, allreduce_fn ar, void* ar_user,
                    bool /*add_prior*/, cudaStream_t stream) {
  if (m.kind == MODEL_LOGISTIC) return launch_glm_t<0>(m, b, ws, ar, ar_user, stream);
  return launch_glm_t<1>(m, b, ws, ar, ar_user, stream);
}

}  // namespace b200
