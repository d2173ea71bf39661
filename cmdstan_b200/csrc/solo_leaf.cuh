// CTA-cooperative execution of the hot part of the NUTS state machine for ONE chain (the latency form of the lock-step engine's
// advance step, used by the kernels that run a single chain to the end on the device: irt.cu, irt_solo_kernel).
//
// advance() (nuts.cuh) gives a chain to one warp: right for thousands of chains side by side, but a straggler chain with a long
// parameter vector then pays ~40 dependent global-memory round trips per vector pass.  Here the 256 threads of a CTA share the
// element-wise work of a tree leaf -- base_nuts::build_tree's base case and the merges that follow it (base_nuts.hpp:253-351),
// expl_leapfrog (expl_leapfrog.hpp:16-32) -- so a pass is one round trip.
//
// Bit-identity with advance() is kept on purpose (a chain may be served by either, depending on how many OTHER chains are still
// running; its draws must not show it): element-wise results are the same operations on the same operands whoever computes them,
// and every reduction is funnelled through warp 0 in advance()'s own order -- the per-element products are staged in shared
// memory, lane l adds elements l, l + 32, ... sequentially, then the same xor butterfly.  Everything outside the leaf path (end
// of a subtree, doublings, transitions, adaptation) is rare and stays with advance() itself, run by warp 0.
// Only the diag_e / unit_e NUTS path (no dense metric, no diagnostics rows, no static HMC).
#pragma once
#include "nuts.cuh"

namespace b200 {

constexpr int SOLO_NT = 256;
enum SoloResult { SOLO_NEED_GRAD = 0, SOLO_GENERIC = 1 };

struct SoloFlags {
  int valid, lc, take_right, result;
};

// lane `lane` of warp 0: sum of terms[lane], terms[lane + 32], ... in index order (advance()'s per-lane order)
__device__ __forceinline__ double solo_lane_sum(const double* terms, int D, int lane) {
  double acc = 0;
#pragma unroll 4
  for (int i = lane; i < D; i += 32) acc += terms[i];
  return acc;
}

// Precondition (CTA-uniform): v.s points to the chain's scalar block in SHARED memory, v.s->state == ST_LEAF, the gradient
// stage's result for z.q is in v.g / *v.zV and visible to this CTA.  red: shared memory, 6 * cfg.ld doubles.
// Returns SOLO_NEED_GRAD (state == ST_LEAF again: the next leaf's gradient is wanted) or SOLO_GENERIC (state == ST_SUBTREE_END:
// the caller lets warp 0 run advance()).  Ends with a block-wide barrier.
__device__ __forceinline__ int solo_leaf_step(const ChainView& v, const SamplerCfg& cfg, double* red, SoloFlags* fl) {
  ChainS* s = v.s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int D = cfg.D, ld = cfg.ld;
  double* p = slot(v, cfg, V_P);
  const double* im = slot(v, cfg, V_IM);
  // ---- leapfrog_end + the leaf's stack entry (ST_LEAF, nuts.cuh)
  const double eps = s->dir * s->eps;
  const double h = 0.5 * eps;
  {
    double* E = slot(v, cfg, V_STACK + 5 * s->sp);
    for (int i = tid; i < D; i += SOLO_NT) {
      const double gi = v.g[i];
      const double pi = p[i] - h * gi;
      p[i] = pi;
      red[i] = pi * (im[i] * pi);
      E[i] = pi;
      E[ld + i] = pi;
      E[2 * ld + i] = pi;
      E[3 * ld + i] = v.q[i];
      E[4 * ld + i] = gi;
    }
  }
  __syncthreads();
  if (warp == 0) {
    const double acc = solo_lane_sum(red, D, lane);
    const double zV = *v.zV;
    const double hraw = 0.5 * warp_sum(acc) + zV;
    double H = hraw;
    if (isnan(H)) H = INFINITY;
    if (lane == 0) {
      const int e = s->sp;
      s->n_grad++;
      s->n_leapfrog++;
      if ((H - s->H0) > 1000.0) s->divergent = 1;
      const double w = s->H0 - H;
      if (w > 0) s->sum_metro += 1;
      else s->sum_metro += dev_exp(w);
      s->st_lsw[e] = w;  // log_sum_exp(-inf, H0 - h)
      s->st_V[e] = zV;
      s->st_H[e] = hraw;
      s->sp = e + 1;
      s->leaf_count++;
      fl->valid = !s->divergent;
      fl->lc = (int)s->leaf_count;
    }
  }
  __syncthreads();
  // ---- pending merges: leaf number i is followed by ctz(i) merges of the two topmost stack entries (merge_top, nuts.cuh)
  while (fl->valid && (fl->lc & 1) == 0) {
    const int li = s->sp - 2, ri = s->sp - 1;
    if (warp == 0) {
      const double lsw_l = s->st_lsw[li], lsw_r = s->st_lsw[ri];
      const double lsw_sub = dev_log_sum_exp(lsw_l, lsw_r);
      bool take_right;
      if (lsw_r > lsw_sub) {
        take_right = true;
      } else {
        const double accept_prob = dev_exp(lsw_r - lsw_sub);
        take_right = warp_uniform01(s->rng, lane) < accept_prob;
      }
      if (lane == 0) {
        fl->take_right = take_right ? 1 : 0;
        s->st_lsw[li] = lsw_sub;
        if (take_right) {
          s->st_V[li] = s->st_V[ri];
          s->st_H[li] = s->st_H[ri];
        }
      }
    }
    __syncthreads();
    {
      double* L = slot(v, cfg, V_STACK + 5 * li);
      const double* R = slot(v, cfg, V_STACK + 5 * ri);
      const bool take_right = fl->take_right != 0;
      for (int i = tid; i < D; i += SOLO_NT) {
        const double m = im[i];
        const double lpb = L[i], lpe = L[ld + i], lr = L[2 * ld + i];
        const double rpb = R[i], rpe = R[ld + i], rr = R[2 * ld + i];
        const double sh_lpb = m * lpb, sh_lpe = m * lpe, sh_rpb = m * rpb, sh_rpe = m * rpe;
        const double rs = lr + rr;   // rho_subtree = rho_init + rho_final
        red[i] = sh_rpe * rs;
        red[ld + i] = sh_lpb * rs;
        const double re = lr + rpb;  // rho_init + p_final_beg
        red[2 * ld + i] = sh_rpb * re;
        red[3 * ld + i] = sh_lpb * re;
        const double rf = rr + lpe;  // rho_final + p_init_end
        red[4 * ld + i] = sh_rpe * rf;
        red[5 * ld + i] = sh_lpe * rf;
        L[2 * ld + i] = rs;
        L[ld + i] = rpe;
        if (take_right) {
          L[3 * ld + i] = R[3 * ld + i];
          L[4 * ld + i] = R[4 * ld + i];
        }
      }
    }
    __syncthreads();
    if (warp == 0) {
      double a1 = solo_lane_sum(red, D, lane), a2 = solo_lane_sum(red + ld, D, lane);
      double b1 = solo_lane_sum(red + 2 * ld, D, lane), b2 = solo_lane_sum(red + 3 * ld, D, lane);
      double c1 = solo_lane_sum(red + 4 * ld, D, lane), c2 = solo_lane_sum(red + 5 * ld, D, lane);
      warp_sum6(a1, a2, b1, b2, c1, c2);
      bool persist = (a1 > 0 && a2 > 0);
      persist &= (b1 > 0 && b2 > 0);
      persist &= (c1 > 0 && c2 > 0);
      if (lane == 0) {
        s->sp = ri;
        fl->valid = persist ? 1 : 0;
        fl->lc >>= 1;
      }
    }
    __syncthreads();
  }
  // ---- next leaf of the same subtree (ST_LEAF_BEGIN: leapfrog_begin), or hand the end of the subtree to advance()
  const bool more = fl->valid && s->leaf_count < s->n_leaves;
  __syncthreads();   // every thread has read the flags and counters before thread 0 changes the state
  if (more) {
    for (int i = tid; i < D; i += SOLO_NT) {
      const double pi = p[i] - h * v.g[i];
      p[i] = pi;
      v.q[i] += eps * (im[i] * pi);
    }
    if (tid == 0) s->state = ST_LEAF;
  } else if (tid == 0) {
    s->sub_valid = fl->valid ? 1 : 0;
    s->state = ST_SUBTREE_END;
  }
  __syncthreads();
  return more ? SOLO_NEED_GRAD : SOLO_GENERIC;
}

}  // namespace b200
