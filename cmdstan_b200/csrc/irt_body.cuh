// The CTA-cooperative log-density + gradient of ONE chain of the hierarchical IRT 2PL model (see irt.cu for the model, the
// work distribution and the reference citations), written as three pieces so that two kernels can share it bit for bit:
//
//   irt_prologue   item parameters exp(u_a), b of the chain into registers (one CTA per SM) or shared memory (two per SM)
//   irt_block      the responses of ONE person block (persons are cut into IRT_NB = 8 contiguous blocks): theta gradients written,
//                  per-warp partials of the item gradients and of the log-likelihood returned
//   irt_finish     item-level priors, hierarchical scalars, the potential
//
// The summation order is defined over the blocks ("canonical order"): a lane sums its responses of block b sequentially, lanes
// are combined by fixed shuffle trees inside a warp, a warp's block partials are added over b = 0..7 in order, warps are added
// in warp order.  The throughput kernel (irt_grad_kernel: one CTA per chain) walks the eight blocks itself; the latency kernel
// (irt_solo_kernel: a cluster of eight CTAs per chain, irt.cu) gives block b to CTA b and CTA 0 adds the partials over b through
// distributed shared memory in the same order -- a chain's gradient is the same bits whichever kernel serves a leapfrog, so
// which one does (a function of how many OTHER chains are still running) never shows in its draws (the parallel-match
// invariant, ST/unit/services/sample/hmc_nuts_diag_e_adapt_parallel_match_test.cpp:73-153).  The translation unit is built
// with --fmad=false and every fusion is an explicit fma(), so the two instantiations cannot contract differently.
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
constexpr int IRT_MAXJ = IRT_G * IRT_MAX_IPL;
constexpr int IRT_NB = 8;           // person blocks of the canonical summation order (= CTAs per cluster of the latency kernel)

struct IrtShared {
  double ga[IRT_NW][IRT_MAXJ];   // per-warp partials of d lp / d a_j (before the chain rule), d lp / d b_j
  double gb[IRT_NW][IRT_MAXJ];
  double blk[IRT_NW][4];         // latency kernel: this CTA's block partials per warp: product, exponent, sum, sum of squares
  double lp[IRT_NW], s2[IRT_NW];
  double it[4][IRT_NW];
  double a[IRT_MAXJ], b[IRT_MAXJ];  // two-CTAs-per-SM form: exp(u_a), b of the chain
  double exptab[32];
};

// q is read with ld.global.cg instead of through the read-only path when COHERENT: needed when q was written earlier in the
// SAME kernel (the latency kernel advances the chain and evaluates its gradient in one launch).
#define IRT_LDQ(p) (COHERENT ? __ldcg(p) : *(p))

struct IrtChainPtrs {
  const double *theta, *u_a, *bq;
  double u_st, u_sa, mu_b, u_sb, st, sa, sb;
};

template <bool COHERENT>
__device__ __forceinline__ IrtChainPtrs irt_chain_ptrs(const double* __restrict__ q, int I, int J) {
  IrtChainPtrs c;
  c.u_st = IRT_LDQ(q);
  c.theta = q + 1;
  c.u_sa = IRT_LDQ(q + 1 + I);
  c.u_a = q + 2 + I;
  c.mu_b = IRT_LDQ(q + 2 + I + J);
  c.u_sb = IRT_LDQ(q + 3 + I + J);
  c.bq = q + 4 + I + J;
  c.st = exp(c.u_st);
  c.sa = exp(c.u_sa);
  c.sb = exp(c.u_sb);
  return c;
}

// item parameters of the chain: registers a[], b[] (OCC == 1) or sm.a / sm.b (OCC >= 2); the exp table; zeroes the running
// per-warp item partials.  Ends with a block-wide barrier.
template <int IPL, int OCC, bool COHERENT>
__device__ __forceinline__ void irt_prologue(IrtShared& sm, const IrtChainPtrs& c, int J, double (&a)[OCC >= 2 ? 1 : IPL], double (&b)[OCC >= 2 ? 1 : IPL]) {
  const int tid = threadIdx.x, sub = tid & (IRT_G - 1);
  exp_table_load(sm.exptab, tid);
  if (tid < IRT_MAXJ) {
    if (OCC >= 2) {
      sm.a[tid] = tid < J ? exp(IRT_LDQ(c.u_a + tid)) : 0.0;   // a = 0 for padding items: eta = 0 exactly (logit_acc_value)
      sm.b[tid] = tid < J ? IRT_LDQ(c.bq + tid) : 0.0;
    }
#pragma unroll
    for (int w = 0; w < IRT_NW; w++) {
      sm.ga[w][tid] = 0.0;
      sm.gb[w][tid] = 0.0;
    }
  }
  if (OCC < 2) {
#pragma unroll
    for (int t = 0; t < IPL; t++) {
      const int j = sub + IRT_G * t;
      a[t] = j < J ? exp(IRT_LDQ(c.u_a + j)) : 0.0;
      b[t] = j < J ? IRT_LDQ(c.bq + j) : 0.0;
    }
  }
  __syncthreads();
}

// person block `blk` of the chain: writes g[1 + i] for its persons; returns (valid in every lane of a warp) the warp's
// partials: va[t], vb[t] for item j = sub + 8 t, and the log-likelihood in product form (w_prod in [1, 2), w_expo, w_sum) plus
// the sum of squared standardised abilities w_s2.
template <int IPL, int OCC, bool COHERENT>
__device__ __forceinline__ void irt_block(const IrtShared& sm, const IrtChainPtrs& c, const unsigned short* __restrict__ ypk, int I, int J, int blk,
                                          const double (&a)[OCC >= 2 ? 1 : IPL], const double (&b)[OCC >= 2 ? 1 : IPL], double* __restrict__ g,
                                          double (&va)[IPL], double (&vb)[IPL], double& w_prod, int& w_expo, double& w_sum, double& w_s2) {
  const int tid = threadIdx.x;
  const int sub = tid & (IRT_G - 1);             // item lane inside the group
  const int grp = tid / IRT_G;                   // 0..31
  constexpr int NGRP = IRT_NT / IRT_G;
  const int PB = (I + IRT_NB - 1) / IRT_NB;
  const int base = blk * PB;
  const int end = base + PB < I ? base + PB : I;
#pragma unroll
  for (int t = 0; t < IPL; t++) va[t] = vb[t] = 0.0;
  // log-likelihood of this thread's responses in product form (fastmath.cuh): one multiplication per response
  LogitAcc lacc;
  logit_acc_init(lacc);
  double s2 = 0.0;
  int since_fold = 0, my_persons = 0;
  for (int i0 = base; i0 < end; i0 += NGRP) {  // warp-uniform trip count: the shuffles below need every lane
    const int i = i0 + grp;
    const bool person = i < end;
    const double th = person ? IRT_LDQ(c.theta + i) : 0.0;
    const unsigned ybits = person ? (unsigned)ypk[(size_t)i * IRT_G + sub] : 0u;   // this lane's responses of the person, bit t = item sub + 8 t
    double gth = 0.0;
    if (person) {
      my_persons++;
#pragma unroll
      for (int t = 0; t < IPL; t++) {
        const int j = sub + IRT_G * t;
        const int negmask = (int)((~ybits << (31 - t)) & 0x80000000u);   // sign bit of 2y - 1
        const double at = OCC >= 2 ? sm.a[j] : a[t], bt = OCC >= 2 ? sm.b[j] : b[t];
        const double diff = th - bt;
        const double d = logit_term(lacc, at * diff, negmask, sm.exptab);   // padding items: at == 0 => eta == 0 (see logit_acc_value)
        const double da = d * at;
        gth += da;
        vb[t] -= da;
        va[t] = fma(d, diff, va[t]);
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
      const double z = th / c.st;  // theta ~ normal(0, sigma_theta)
      s2 = fma(z, z, s2);
      g[1 + i] = -(gth - z / c.st);
    }
  }
  // ---- item partials over the 4 groups of the warp (lanes with equal `sub`)
#pragma unroll
  for (int t = 0; t < IPL; t++) {
    va[t] += __shfl_xor_sync(0xffffffffu, va[t], 8);
    vb[t] += __shfl_xor_sync(0xffffffffu, vb[t], 8);
    va[t] += __shfl_xor_sync(0xffffffffu, va[t], 16);
    vb[t] += __shfl_xor_sync(0xffffffffu, vb[t], 16);
  }
  // ---- log-likelihood of the warp's responses of this block: the lanes' products multiplied, exponents and sums added
  logit_acc_fold(lacc);
  int n_pad = 0;   // padding item slots (j >= J) this thread evaluated: one per slot and person
#pragma unroll
  for (int t = 0; t < IPL; t++) n_pad += (sub + IRT_G * t >= J) ? my_persons : 0;
  double pr = lacc.prod, sm_ = lacc.sum;
  int ex = lacc.expo - n_pad;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    pr *= __shfl_xor_sync(0xffffffffu, pr, o);   // 32 factors in [1, 2): below 2^32
    ex += __shfl_xor_sync(0xffffffffu, ex, o);
    sm_ += __shfl_xor_sync(0xffffffffu, sm_, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  LogitAcc wacc;
  wacc.prod = pr;
  wacc.sum = sm_;
  wacc.expo = ex;
  logit_acc_fold(wacc);
  w_prod = wacc.prod;
  w_expo = wacc.expo;
  w_sum = wacc.sum;
  w_s2 = s2;
}

// running per-warp totals over the person blocks, in block order
struct IrtWarpTotals {
  double P = 1.0, S = 0.0, S2 = 0.0;
  int E = 0;
  __device__ __forceinline__ void add(double w_prod, int w_expo, double w_sum, double w_s2) {
    LogitAcc t;
    t.prod = P * w_prod;
    t.expo = E + w_expo;
    t.sum = S + w_sum;
    logit_acc_fold(t);
    P = t.prod;
    E = t.expo;
    S = t.sum;
    S2 += w_s2;
  }
  __device__ __forceinline__ double lp() const { return -(S + fma((double)E, 0.69314718055994530942, log(P))); }
};

// item-level terms, hierarchical scalars, potential.  In: sm.ga / sm.gb [w][j] (each warp's total over the blocks), sm.lp[w],
// sm.s2[w]; must be preceded by a block-wide barrier.  Ends with every shared array consumed before its trailing barrier... the
// caller synchronises before reusing `sm`.
template <bool COHERENT>
__device__ __forceinline__ void irt_finish(IrtShared& sm, const IrtChainPtrs& c, int I, int J, double* __restrict__ g, double* __restrict__ zVc) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // ---- item-level terms: thread j sums the per-warp partials in warp order, adds the priors
  double s2a = 0.0, sua = 0.0, s2b = 0.0, gmub = 0.0;
  if (tid < IRT_MAXJ) {
    const int j = tid;
    if (j < J) {
      double gaj = 0.0, gbj = 0.0;
      for (int w = 0; w < IRT_NW; w++) {
        gaj += sm.ga[w][j];
        gbj += sm.gb[w][j];
      }
      const double ua = IRT_LDQ(c.u_a + j);
      const double aj = exp(ua);
      const double za = ua / c.sa;                     // a ~ lognormal(0, sigma_a), a = exp(ua), Jacobian +ua
      const double zb = (IRT_LDQ(c.bq + j) - c.mu_b) / c.sb;     // b ~ normal(mu_b, sigma_b)
      s2a = za * za;
      sua = ua;
      s2b = zb * zb;
      gmub = zb / c.sb;
      g[2 + I + j] = -(gaj * aj + (-za / c.sa - 1.0 + 1.0));
      g[4 + I + J + j] = -(gbj - zb / c.sb);
    }
  }
  // block-wide fixed-order sums of the four item statistics (threads 0..127 hold them): warp sums, then warp order
  s2a = warp_sum(s2a);
  sua = warp_sum(sua);
  s2b = warp_sum(s2b);
  gmub = warp_sum(gmub);
  if (lane == 0) {
    sm.it[0][warp] = s2a;
    sm.it[1][warp] = sua;
    sm.it[2][warp] = s2b;
    sm.it[3][warp] = gmub;
  }
  __syncthreads();
  if (tid == 0) {
    const double NEG_LOG_SQRT_TWO_PI = -0.91893853320467274178;
    const double u_st = c.u_st, u_sa = c.u_sa, u_sb = c.u_sb, st = c.st, sa = c.sa, sb = c.sb, mu_b = c.mu_b;
    double lpl = 0.0, S2 = 0.0, S2A = 0.0, SUA = 0.0, S2B = 0.0, GMUB = 0.0;
    for (int w = 0; w < IRT_NW; w++) {
      lpl += sm.lp[w];
      S2 += sm.s2[w];
      S2A += sm.it[0][w];
      SUA += sm.it[1][w];
      S2B += sm.it[2][w];
      GMUB += sm.it[3][w];
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

// the whole chain by one CTA (throughput form): the eight person blocks in order.  q: the chain's unconstrained parameters [D];
// g: -grad lp out [D]; zVc: -lp out.  Every shared array is consumed before the function returns except sm.it / sm.lp / sm.s2,
// read by thread 0 after the last barrier: a caller that loops must synchronise before the next call.
template <int IPL, int OCC, bool COHERENT = false>
__device__ __forceinline__ void irt_chain_body(IrtShared& sm, const unsigned short* __restrict__ y, int I, int J, int D, const double* __restrict__ q,
                                               double* __restrict__ g, double* __restrict__ zVc) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, sub = tid & (IRT_G - 1);
  const IrtChainPtrs c = irt_chain_ptrs<COHERENT>(q, I, J);
  double a[OCC >= 2 ? 1 : IPL], b[OCC >= 2 ? 1 : IPL];
  irt_prologue<IPL, OCC, COHERENT>(sm, c, J, a, b);
  IrtWarpTotals tot;
  for (int blk = 0; blk < IRT_NB; blk++) {
    double va[IPL], vb[IPL], w_prod, w_sum, w_s2;
    int w_expo;
    irt_block<IPL, OCC, COHERENT>(sm, c, y, I, J, blk, a, b, g, va, vb, w_prod, w_expo, w_sum, w_s2);
    if (lane < IRT_G) {
#pragma unroll
      for (int t = 0; t < IPL; t++) {
        sm.ga[warp][sub + IRT_G * t] += va[t];
        sm.gb[warp][sub + IRT_G * t] += vb[t];
      }
    }
    tot.add(w_prod, w_expo, w_sum, w_s2);
  }
  if (lane == 0) {
    sm.lp[warp] = tot.lp();
    sm.s2[warp] = tot.S2;
  }
  __syncthreads();
  irt_finish<COHERENT>(sm, c, I, J, g, zVc);
}
#undef IRT_LDQ

}  // namespace b200
