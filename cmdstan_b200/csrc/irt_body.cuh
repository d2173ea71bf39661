// The CTA-cooperative log-density + gradient of ONE chain of the hierarchical IRT 2PL model (see irt.cu for the model, the
// work distribution and the reference citations).  Shared by the cross-chain gradient kernel (irt.cu) and the experimental
// straggler kernel (irt_tail.cu), which calls it in a loop: every shared array is consumed before the trailing
// __syncthreads of the caller's loop.
#pragma once
#include "fastmath.cuh"
#include "kernels.h"

#ifdef B200_SIMT_EMUL  // CPU test-suite only (tests/host_emul/simt): the kernels run on host threads
#define IRT_KERNEL_LAUNCH(kernel, grid, stream, ...) simt::launch(grid, IRT_NT, 0, [&] { kernel(__VA_ARGS__); })
#else
#define IRT_KERNEL_LAUNCH(kernel, grid, stream, ...) kernel<<<grid, IRT_NT, 0, stream>>>(__VA_ARGS__)
#endif

namespace b200 {

constexpr int IRT_NT = 256;
constexpr int IRT_NW = IRT_NT / 32;
constexpr int IRT_G = 8;            // lanes per person group
constexpr int IRT_MAX_IPL = 16;     // items per lane


// q: the chain's unconstrained parameters [D]; g: -grad lp out [D]; zVc: -lp out.
// COHERENT: q is read with ld.global.cg instead of through the read-only path -- needed when q was written earlier in the
// SAME kernel (the straggler kernel advances the chain and evaluates its gradient in one launch).
#define IRT_LDQ(p) (COHERENT ? __ldcg(p) : *(p))

template <int IPL, int OCC, bool COHERENT = false>
__device__ __forceinline__ void irt_chain_body(const signed char* __restrict__ y, int I, int J, int D, const double* __restrict__ q,
                                               double* __restrict__ g, double* __restrict__ zVc) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = lane & (IRT_G - 1);             // item lane inside the group
  const int grp = tid / IRT_G;                    // 0..31
  constexpr int NGRP = IRT_NT / IRT_G;
  __shared__ double s_ga[IRT_NW][IRT_G * IRT_MAX_IPL];
  __shared__ double s_gb[IRT_NW][IRT_G * IRT_MAX_IPL];
  __shared__ double s_lp[IRT_NW];
  __shared__ double s_s2[IRT_NW];

  const double u_st = IRT_LDQ(q);
  const double* theta = q + 1;
  const double u_sa = IRT_LDQ(q + 1 + I);
  const double* u_a = q + 2 + I;
  const double mu_b = IRT_LDQ(q + 2 + I + J);
  const double u_sb = IRT_LDQ(q + 3 + I + J);
  const double* bq = q + 4 + I + J;
  const double st = exp(u_st), sa = exp(u_sa), sb = exp(u_sb);

  __shared__ double s_a[OCC == 2 ? IRT_G * IRT_MAX_IPL : 1], s_b[OCC == 2 ? IRT_G * IRT_MAX_IPL : 1];
  __shared__ double s_exptab[32];
  double a[OCC == 2 ? 1 : IPL], b[OCC == 2 ? 1 : IPL], ga[IPL], gb[IPL];
  exp_table_load(s_exptab, tid);
  if (OCC == 2) {
    if (tid < IRT_G * IRT_MAX_IPL) {
      s_a[tid] = tid < J ? exp(IRT_LDQ(u_a + tid)) : 0.0;   // a = 0 for padding items: no gradient contribution by construction
      s_b[tid] = tid < J ? IRT_LDQ(bq + tid) : 0.0;
    }
  }
  __syncthreads();
#pragma unroll
  for (int t = 0; t < IPL; t++) {
    const int j = sub + IRT_G * t;
    if (OCC != 2) {
      a[t] = j < J ? exp(IRT_LDQ(u_a + j)) : 0.0;
      b[t] = j < J ? IRT_LDQ(bq + j) : 0.0;
    }
    ga[t] = 0.0;
    gb[t] = 0.0;
  }
  // log-likelihood of this thread's responses in product form (fastmath.cuh): one multiplication per response, one log at the end
  LogitAcc lacc;
  logit_acc_init(lacc);
  double s2 = 0.0;
  int since_fold = 0;
  for (int i0 = 0; i0 < I; i0 += NGRP) {  // warp-uniform trip count: the shuffles below need every lane
    const int i = i0 + grp;
    const bool person = i < I;
    const double th = person ? IRT_LDQ(theta + i) : 0.0;
    const signed char* yrow = y + (size_t)(person ? i : 0) * J;
    double gth = 0.0;
    if (person) {
#pragma unroll
    for (int t = 0; t < IPL; t++) {
      const int j = sub + IRT_G * t;
      const int yv = j < J ? yrow[j] : 0;
      const double sgn = __hiloint2double((int)(0xbff00000u ^ ((unsigned)yv << 31)), 0);   // 2y - 1
      const double at = OCC == 2 ? s_a[j] : a[t], bt = OCC == 2 ? s_b[j] : b[t];
      const double diff = th - bt;
      const double d = logit_term(lacc, at * diff, sgn, s_exptab);   // padding items: at == 0 => eta == 0 (see logit_acc_value)
      const double da = d * at;
      gth += da;
      gb[t] -= da;
      ga[t] += d * diff;
    }
    }
    since_fold += IPL;
    if (since_fold > 800) {  // the running product gains at most one binade per response
      logit_acc_fold(lacc);
      since_fold = 0;
    }
    gth += __shfl_xor_sync(0xffffffffu, gth, 1);
    gth += __shfl_xor_sync(0xffffffffu, gth, 2);
    gth += __shfl_xor_sync(0xffffffffu, gth, 4);
    if (sub == 0 && person) {
      const double z = th / st;  // theta ~ normal(0, sigma_theta)
      s2 += z * z;
      g[1 + i] = -(gth - z / st);
    }
  }
  // ---- item partials: over the 4 groups of the warp (lanes with equal `sub`), then over warps in warp order
#pragma unroll
  for (int t = 0; t < IPL; t++) {
    double va = ga[t], vb = gb[t];
    va += __shfl_xor_sync(0xffffffffu, va, 8);
    vb += __shfl_xor_sync(0xffffffffu, vb, 8);
    va += __shfl_xor_sync(0xffffffffu, va, 16);
    vb += __shfl_xor_sync(0xffffffffu, vb, 16);
    if (lane < IRT_G) {
      s_ga[warp][sub + IRT_G * t] = va;
      s_gb[warp][sub + IRT_G * t] = vb;
    }
  }
  logit_acc_fold(lacc);
  int n_pad = 0;   // padding item slots (j >= J) this thread evaluated: one per slot and person
  {
    int my_persons = 0;
    for (int i0 = 0; i0 < I; i0 += NGRP) my_persons += (i0 + grp < I) ? 1 : 0;
#pragma unroll
    for (int t = 0; t < IPL; t++) n_pad += (sub + IRT_G * t >= J) ? my_persons : 0;
  }
  double lp = logit_acc_value(lacc, n_pad);
  lp = warp_sum(lp);
  s2 = warp_sum(s2);
  if (lane == 0) {
    s_lp[warp] = lp;
    s_s2[warp] = s2;
  }
  __syncthreads();
  // ---- item-level terms: thread j sums the per-warp partials in warp order, adds the priors
  double s2a = 0.0, sua = 0.0, s2b = 0.0, gmub = 0.0;
  if (tid < IRT_G * IRT_MAX_IPL) {
    const int j = tid;
    if (j < J) {
      double gaj = 0.0, gbj = 0.0;
      for (int w = 0; w < IRT_NW; w++) {
        gaj += s_ga[w][j];
        gbj += s_gb[w][j];
      }
      const double ua = IRT_LDQ(u_a + j);
      const double aj = exp(ua);
      const double za = ua / sa;                 // a ~ lognormal(0, sigma_a), a = exp(ua), Jacobian +ua
      const double zb = (IRT_LDQ(bq + j) - mu_b) / sb;     // b ~ normal(mu_b, sigma_b)
      s2a = za * za;
      sua = ua;
      s2b = zb * zb;
      gmub = zb / sb;
      g[2 + I + j] = -(gaj * aj + (-za / sa - 1.0 + 1.0));
      g[4 + I + J + j] = -(gbj - zb / sb);
    }
  }
  // block-wide fixed-order sums of the four item statistics (threads 0..127 hold them): warp sums, then warp order
  __shared__ double s_it[4][IRT_NW];
  s2a = warp_sum(s2a);
  sua = warp_sum(sua);
  s2b = warp_sum(s2b);
  gmub = warp_sum(gmub);
  if (lane == 0) {
    s_it[0][warp] = s2a;
    s_it[1][warp] = sua;
    s_it[2][warp] = s2b;
    s_it[3][warp] = gmub;
  }
  __syncthreads();
  if (tid == 0) {
    const double NEG_LOG_SQRT_TWO_PI = -0.91893853320467274178;
    double lpl = 0.0, S2 = 0.0, S2A = 0.0, SUA = 0.0, S2B = 0.0, GMUB = 0.0;
    for (int w = 0; w < IRT_NW; w++) {
      lpl += s_lp[w];
      S2 += s_s2[w];
      S2A += s_it[0][w];
      SUA += s_it[1][w];
      S2B += s_it[2][w];
      GMUB += s_it[3][w];
    }
    double total = u_st + u_sa + u_sb;  // Jacobians of the three scalar lower bounds
    bool ok = st > 0 && isfinite(st) && sa > 0 && isfinite(sa) && sb > 0 && isfinite(sb);
    double h = st / 2.0;
    total += -log1p(h * h);
    double g_st = -2.0 * h * h / (1.0 + h * h);
    total += -0.5 * S2 - I * u_st;
    g_st += S2 - I;
    h = sa / 2.0;
    total += -log1p(h * h);
    double g_sa = -2.0 * h * h / (1.0 + h * h);
    total += J * NEG_LOG_SQRT_TWO_PI - 0.5 * S2A - J * u_sa - SUA + SUA;
    g_sa += S2A - J;
    const double zmb = mu_b / 5.0;
    total += -0.5 * zmb * zmb;
    const double g_mub = -zmb / 5.0 + GMUB;
    h = sb / 2.0;
    total += -log1p(h * h);
    double g_sb = -2.0 * h * h / (1.0 + h * h);
    total += -0.5 * S2B - J * u_sb;
    g_sb += S2B - J;
    total += lpl;
    g[0] = -(g_st + 1.0);
    g[1 + I] = -(g_sa + 1.0);
    g[2 + I + J] = -g_mub;
    g[3 + I + J] = -(g_sb + 1.0);
    *zVc = ok ? -total : INFINITY;  // check_positive_finite on a scale => domain_error => V = +inf
  }
}
#undef IRT_LDQ

}  // namespace b200
