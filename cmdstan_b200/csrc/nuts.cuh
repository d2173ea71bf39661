// The per-chain NUTS state machine: adaptive multinomial NUTS with a diagonal Euclidean metric, exactly the
// semantics (and random-number consumption order) of
//
//   services::sample::hmc_nuts_diag_e_adapt   S/services/sample/hmc_nuts_diag_e_adapt.hpp:58-112  (per-chain set-up)
//   util::initialize                          S/services/util/initialize.hpp:70-260
//   util::run_adaptive_sampler                S/services/util/run_adaptive_sampler.hpp:47-103
//   adapt_diag_e_nuts::transition             S/mcmc/hmc/nuts/adapt_diag_e_nuts.hpp:25-48
//   base_nuts::transition / build_tree        S/mcmc/hmc/nuts/base_nuts.hpp:78-204, :247-352, criterion :222-226
//   base_hmc::init_stepsize / sample_stepsize S/mcmc/hmc/base_hmc.hpp:93-161, :195-200
//   expl_leapfrog                             S/mcmc/hmc/integrators/expl_leapfrog.hpp:16-32
//   diag_e_metric                             S/mcmc/hmc/hamiltonians/diag_e_metric.hpp:20-50
//   dense_e_metric / covar_adaptation         S/mcmc/hmc/hamiltonians/dense_e_metric.hpp:21-58, S/mcmc/covar_adaptation.hpp:17-45,
//                                             M/prim/fun/welford_covar_estimator.hpp:10-50 (metric = dense_e, cfg.dense)
//   stepsize_adaptation / var_adaptation      S/mcmc/stepsize_adaptation.hpp:49-73, S/mcmc/var_adaptation.hpp:17-46,
//   windowed_adaptation / welford             S/mcmc/windowed_adaptation.hpp:23-110, M/prim/fun/welford_var_estimator.hpp
//
// but formulated for the GPU rather than translated:
//
//  * one chain = one warp; vectors are strided over lanes, scalars are warp-uniform, reductions are shuffles;
//  * build_tree's recursion is replaced by an iterative binary-counter merge: leaves are produced one at a time,
//    leaf number i is followed by ctz(i+1) merges of the two topmost stack entries, each merge doing the reference's
//    operations in the reference's order (multinomial pick -> rho sum -> three U-turn criteria).  The stack holds at
//    most `depth` entries of (p_beg, p_end, rho, proposal q/g/V/H); sharp momenta are recomputed (M^-1 o p) instead of
//    stored.  The post-order of merges is the recursion's post-order, so the RNG stream is consumed identically;
//  * the whole sampler is a *resumable* state machine whose only suspension point is "log density and gradient at
//    z.q wanted".  The fused kernel (small models) evaluates the model inline and loops; the lock-step engine
//    (GLM / IRT) suspends every chain, runs one cross-chain gradient kernel for all of them, and resumes;
//  * (V, g) of the current sample are cached, so the reference's extra gradient evaluation at the start of every
//    transition and of every init_stepsize probe (base_hamiltonian::init, same q => same value) is not re-executed.
#pragma once
#include "models.cuh"
#include "rng.cuh"

// Unroll factor of the per-lane loops over a chain's vector elements (i = lane, lane + 32, ...).  1 (default) keeps the
// state machine small: the fused chain kernels of the small models are instruction-cache bound.  The lock-step engine's
// advance kernel for long vectors (advance_big.cu, D >= 64) is built with 8: one warp per chain is latency-bound on its
// dependent global loads (profiles r01e: 895 us per step at D = 1204), and unrolling puts 8 elements per lane in flight.
// Unrolling does not reorder any floating-point sum.
#ifndef B200_VEC_UNROLL
#define B200_VEC_UNROLL 1
#endif
#define B200_PRAGMA_(x) _Pragma(#x)
#define B200_PRAGMA(x) B200_PRAGMA_(x)
#define B200_VEC_LOOP B200_PRAGMA(unroll B200_VEC_UNROLL)
// the six U-turn dot products of a merge: one interleaved butterfly (5 dependent shuffle rounds) in the latency-bound long-vector
// build; six calls of the one shared out-of-line warp_sum in the instruction-cache-bound default build (r01l / r01m: the
// inlined form cost the c2 fused kernel ~4 %).  Same bits either way.
#if B200_VEC_UNROLL > 1
#define B200_SUM6(a, b, c, d, e, f) warp_sum6(a, b, c, d, e, f)
#else
#define B200_SUM6(a, b, c, d, e, f) \
  do {                              \
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c); d = warp_sum(d); e = warp_sum(e); f = warp_sum(f); \
  } while (0)
#endif

// The dense-metric code paths are compiled only where B200_WITH_DENSE is 1 (dense.cu, the host emulation): the fused chain
// kernels of the small models are instruction-cache bound, and carrying the dense branches cost c2 15 % (profiles r01i).
#ifndef B200_WITH_DENSE
#define B200_WITH_DENSE 1
#endif
#define B200_DENSE(cfg) (B200_WITH_DENSE && (cfg).dense)
// the diagnostics bookkeeping (momentum of every proposal) is compiled with the dense paths: runs with store_diagnostics use
// the kernels of dense.cu, the default kernels stay small (r01l: carrying these branches cost the c2 fused kernel 7 %)
#define B200_DIAG(cfg) (B200_WITH_DENSE && (cfg).diag)
// static HMC (S/mcmc/hmc/static/base_static_hmc.hpp) likewise lives only in the full-featured kernels of dense.cu
#define B200_STATIC(cfg) (B200_WITH_DENSE && (cfg).static_hmc)

namespace b200 {

constexpr int kMaxDepthCap = 16;  // stack scalars are sized for trees of up to 2^16 leapfrogs

// vector slots of a chain (each `ld` doubles).  z.q and z.g live in separate dense [chain][ld] arrays because they
// are the operands / results of the cross-chain gradient kernels.
enum VecSlot {
  V_P = 0,     // z.p : momentum of the working point
  V_CQ,        // current sample q   (start of the transition, z_sample during it, z_init of init_stepsize)
  V_CG,        // gradient (of the potential) at V_CQ
  V_FQ, V_FP, V_FG,   // z_fwd
  V_BQ, V_BP, V_BG,   // z_bck
  V_PFB,       // p_fwd_bck
  V_PBF,       // p_bck_fwd
  V_RHO,       // rho of the whole trajectory
  V_IM,        // diagonal of the inverse metric
  V_WM, V_WM2, // Welford accumulators
  V_STACK      // + 5*level + {0: p_beg, 1: p_end, 2: rho, 3: proposal q, 4: proposal g}
};
// dense metric: 4 more scratch vectors after the tree stack (sharp momenta M^-1 p are recomputed where the reference stores them)
// diagnostics (cfg.diag): 1 + max_depth further vectors -- the momentum of the current sample and of each level's proposal
B200_HD int num_vec_slots(int max_depth, int dense = 0, int diag = 0) {
  return V_STACK + 5 * max_depth + (dense ? 4 : 0) + (diag ? 1 + max_depth : 0);
}

enum ChainState {
  ST_START = 0,
  ST_INIT_DRAW,
  ST_INIT_CHECK,      // waits for gradient
  ST_ISS_BEGIN,
  ST_ISS_FIRST,       // waits
  ST_ISS_LOOP_BEGIN,
  ST_ISS_LOOP,        // waits
  ST_ISS_END,
  ST_ITER_BEGIN,      // also the idle state between b200nuts_run calls
  ST_DOUBLING_BEGIN,
  ST_LEAF_BEGIN,
  ST_LEAF,            // waits
  ST_SUBTREE_END,     // the subtree of this doubling is complete (or was abandoned): extend the trajectory, next doubling or end
  ST_STATIC_STEP_BEGIN,  // static HMC: one of the L leapfrogs of a transition
  ST_STATIC_STEP,        // waits
  ST_STATIC_END,
  ST_TRANS_END,
  ST_ITER_NEXT,
  ST_DONE             // chain stopped with an error code
};

enum AdvanceResult { ADV_NEED_GRAD = 0, ADV_IDLE = 1 };

// run-wide configuration (same for every chain; passed to kernels by value)
struct SamplerCfg {
  int D, ld, max_depth;
  int num_warmup, num_samples, num_thin, save_warmup;
  int adapt_engaged, user_inits;
  int static_hmc;        // 1: static HMC (base_static_hmc.hpp:41-70): L = max(1, (int)(int_time / nominal stepsize)) leapfrogs and a
                         //    Metropolis test per transition; rows: lp__, accept_stat__, stepsize__, L, 1, 0, energy__, q
  double int_time;
  int diag;              // 1: every saved row is [7 + 3 D]: the parameters are followed by the sample's momentum and by the gradient
                         //    of the potential (what the reference sends to the diagnostic writer, mcmc_writer.hpp:156-176)
  int dense;             // 1: dense_e metric (per-chain D x D inverse metric, covariance windows); 0: diag_e / unit_e
  double stepsize, stepsize_jitter, delta, gamma, kappa, t0, init_radius;
  // resolved window schedule (windowed_adaptation::set_window_params)
  unsigned w_num_warmup, w_init_buffer, w_term_buffer, w_base_window, w_next0, w_size0;
  int ncol;              // 7 + D
  long long rows_per_chain;
};

// per-chain scalar state (global memory in the lock-step engine, shared memory in the fused kernel)
struct ChainS {
  unsigned long long rng[kRngWords];
  int state, err, iter, iter_end;
  int init_tries, iss_after, iss_dir, dir;
  int init_partial, st_L;   // init_partial: some initial values were left to the random context (up to 100 tries, like no inits)
  int st_i, sub_valid;      // static HMC: L_ (st_L) and the leapfrogs done in this transition (st_i); sub_valid: ST_SUBTREE_END's input
  int depth, n_leapfrog, divergent, sp;
  unsigned leaf_count, n_leaves;
  unsigned win_counter, win_size, win_next;
  int draws_row;
  long long n_grad;        // leapfrog gradients consumed so far
  double eps_nom, eps;
  double curV, curH;       // potential / Hamiltonian of the current sample (curH = energy__ of the last transition)
  double H0, lsw, sum_metro;
  double iss_H0;
  double sa_counter, s_bar, x_bar, mu;
  double wf_n;
  double st_lsw[kMaxDepthCap], st_V[kMaxDepthCap], st_H[kMaxDepthCap];
};

struct ChainView {
  ChainS* s;
  double* q;      // z.q   [ld]
  double* g;      // z.g   [ld]   (gradient of the potential = -grad lp, as stored by the gradient stage)
  double* zV;     // z.V          (potential = -lp, as stored by the gradient stage)
  double* v;      // vector slab, slot k at v + k*ld
  double* draws;  // this chain's rows [rows_per_chain][ncol] or nullptr
  const double* user_q0;  // [ld] or nullptr
  double* mat = nullptr;  // dense metric only: three D x D column-major matrices (element (i,j) at j*D + i, like
                          // Eigen::MatrixXd): [0] inverse metric M^-1, [1] its lower Cholesky factor L, [2] Welford m2
};

B200_DEV double* slot(const ChainView& c, const SamplerCfg& cfg, int k) { return c.v + (size_t)k * cfg.ld; }

// "not given" marker of user inits: a quiet NaN whose low payload bits carry the parameter's transform (see ST_INIT_DRAW)
B200_HD int init_nan_code(double v) {
  unsigned long long bits;
  memcpy(&bits, &v, sizeof bits);
  return (int)(bits & 3ull);
}
B200_HD double init_nan_with_code(int code) {
  const unsigned long long bits = 0x7ff8000000000000ull | (unsigned long long)(code & 3);
  double v;
  memcpy(&v, &bits, sizeof v);
  return v;
}
B200_DEV double init_logit(double th) { return dev_log(th / (1.0 - th)); }  // M/prim/fun/logit.hpp: log(u / (1 - u))

B200_DEV double* scratch(const ChainView& c, const SamplerCfg& cfg, int k) { return slot(c, cfg, V_STACK + 5 * cfg.max_depth + k); }
// diagnostics slots: k = 0 momentum of the current sample, k = 1 + level momentum of that level's proposal
B200_DEV double* diag_slot(const ChainView& c, const SamplerCfg& cfg, int k) {
  return slot(c, cfg, V_STACK + 5 * cfg.max_depth + (cfg.dense ? 4 : 0) + k);
}

// ---- dense metric -----------------------------------------------------------------------------------------------------
// out = M^-1 p  (dense_e_metric::dtau_dp, dense_e_metric.hpp:38).  Lane l owns rows l, l+32, ...; the matrix is column-major
// so the 32 lanes read consecutive addresses; each row is accumulated over the columns in index order (the order of Eigen's
// column-major gemv kernel).
B200_DEV_CALL void dense_sharp(const double* M, const double* p, double* out, int D, int lane) {
  warp_sync();  // p was written by its owner lanes
  B200_VEC_LOOP for (int i = lane; i < D; i += kWarp) {
    double acc = 0;
    _Pragma("unroll 4") for (int j = 0; j < D; j++) acc += M[(size_t)j * D + i] * p[j];
    out[i] = acc;
  }
}
// lower Cholesky factor L of M^-1 into mat[1] (the reference factorises inside every sample_p, dense_e_metric.hpp:55
// `inv_e_metric_.llt()`; the factor only changes with the metric, so it is computed once per metric).  Reads the lower
// triangle only, like Eigen::LLT<Lower>.  Returns false when the matrix is not positive definite.
B200_DEV_CALL bool dense_factor(double* mat, int D, int lane) {
  const double* A = mat;
  double* L = mat + (size_t)D * D;
  B200_VEC_LOOP for (int i = lane; i < D; i += kWarp)
    for (int j = 0; j < D; j++) L[(size_t)j * D + i] = 0.0;
  warp_sync();
  for (int j = 0; j < D; j++) {
    double d = A[(size_t)j * D + j];  // every lane forms the pivot itself (same values, same order)
    for (int k = 0; k < j; k++) d -= L[(size_t)k * D + j] * L[(size_t)k * D + j];
    if (!(d > 0)) return false;
    const double ljj = sqrt(d);
    B200_VEC_LOOP for (int i = j + 1 + ((lane - (j + 1)) % kWarp + kWarp) % kWarp; i < D; i += kWarp) {  // rows i > j owned by this lane
      double v = A[(size_t)j * D + i];
      for (int k = 0; k < j; k++) v -= L[(size_t)k * D + i] * L[(size_t)k * D + j];
      L[(size_t)j * D + i] = v / ljj;
    }
    if (lane == j % kWarp) L[(size_t)j * D + j] = ljj;
    warp_sync();
  }
  return true;
}
// p = L^-T u  (dense_e_metric::sample_p, dense_e_metric.hpp:44-56: matrixU().solve(u)): back substitution on L^T
B200_DEV_CALL void dense_solve_upper(const double* L, const double* u, double* p, int D, int lane) {
  warp_sync();
  for (int i = D - 1; i >= 0; i--) {
    double acc = 0;
    B200_VEC_LOOP for (int k = i + 1 + lane; k < D; k += kWarp) acc += L[(size_t)i * D + k] * p[k];
    acc = warp_sum(acc);
    if (lane == i % kWarp) p[i] = (u[i] - acc) / L[(size_t)i * D + i];
    warp_sync();
  }
}

// T(p) = 0.5 p . (M^-1 o p)   (diag_e_metric.hpp:20-22)
B200_DEV double kinetic(const double* p, const double* im, int D, int lane) {
  double acc = 0;
  B200_VEC_LOOP for (int i = lane; i < D; i += kWarp) acc += p[i] * (im[i] * p[i]);
  return 0.5 * warp_sum(acc);
}

// dense_e_metric::T (dense_e_metric.hpp:21-23): 0.5 p^T (M^-1 p)
B200_DEV_CALL double kinetic_dense(const ChainView& c, const SamplerCfg& cfg, const double* p, int lane) {
  double* s0 = scratch(c, cfg, 0);
  dense_sharp(c.mat, p, s0, cfg.D, lane);
  double acc = 0;
  B200_VEC_LOOP for (int i = lane; i < cfg.D; i += kWarp) acc += p[i] * s0[i];
  return 0.5 * warp_sum(acc);
}

// first two thirds of expl_leapfrog::evolve: p -= (eps/2) g ; q += eps (M^-1 o p); then the gradient is wanted
B200_DEV void leapfrog_begin(const ChainView& c, const SamplerCfg& cfg, double eps, int lane) {
  const double h = 0.5 * eps;
  double* p = slot(c, cfg, V_P);
  if (B200_DENSE(cfg)) {
    B200_VEC_LOOP for (int i = lane; i < cfg.D; i += kWarp) p[i] -= h * c.g[i];
    double* s0 = scratch(c, cfg, 0);
    dense_sharp(c.mat, p, s0, cfg.D, lane);
    B200_VEC_LOOP for (int i = lane; i < cfg.D; i += kWarp) c.q[i] += eps * s0[i];
    return;
  }
  const double* im = slot(c, cfg, V_IM);
  B200_VEC_LOOP for (int i = lane; i < cfg.D; i += kWarp) {
    double pi = p[i] - h * c.g[i];
    p[i] = pi;
    c.q[i] += eps * (im[i] * pi);
  }
}
// last third: p -= (eps/2) g(new); returns H = T + V (NaN -> +inf as base_nuts.hpp:259-260 / base_hmc.hpp:118-119)
B200_DEV double leapfrog_end(const ChainView& c, const SamplerCfg& cfg, double eps, int lane, double* h_raw) {
  const double h = 0.5 * eps;
  double* p = slot(c, cfg, V_P);
  const double* im = slot(c, cfg, V_IM);
  double acc = 0;
  double H;
  if (B200_DENSE(cfg)) {
    B200_VEC_LOOP for (int i = lane; i < cfg.D; i += kWarp) p[i] -= h * c.g[i];
    H = kinetic_dense(c, cfg, p, lane) + (*c.zV);
  } else {
    B200_VEC_LOOP for (int i = lane; i < cfg.D; i += kWarp) {
      double pi = p[i] - h * c.g[i];
      p[i] = pi;
      acc += pi * (im[i] * pi);
    }
    H = 0.5 * warp_sum(acc) + (*c.zV);
  }
  *h_raw = H;
  if (isnan(H)) H = INFINITY;
  return H;
}

// z := current sample (q, g, V); fresh momentum; returns H0.  (hamiltonian.sample_p + hamiltonian.init with the cached gradient)
B200_DEV_CALL double refresh_from_current(const ChainView& c, const SamplerCfg& cfg, int lane) {
  double* p = slot(c, cfg, V_P);
  const double* im = slot(c, cfg, V_IM);
  const double* cq = slot(c, cfg, V_CQ);
  const double* cg = slot(c, cfg, V_CG);
  if (B200_DENSE(cfg)) {  // u ~ N(0, I) in index order, p = L^-T u
    double* u = scratch(c, cfg, 1);
    warp_fill(c.s->rng, lane, u, cfg.D, 0, nullptr, 0.0);
    dense_solve_upper(c.mat + (size_t)cfg.D * cfg.D, u, p, cfg.D, lane);
  } else {
    warp_fill(c.s->rng, lane, p, cfg.D, 0, im, 0.0);
  }
  B200_VEC_LOOP for (int i = lane; i < cfg.D; i += kWarp) {
    c.q[i] = cq[i];
    c.g[i] = cg[i];
  }
  (*c.zV) = c.s->curV;
  if (B200_DENSE(cfg)) return kinetic_dense(c, cfg, p, lane) + c.s->curV;
  return kinetic(p, im, cfg.D, lane) + c.s->curV;
}

// Merge of the two topmost stack entries L (level sp-2) and R (sp-1) into L: base_nuts.hpp:313-351.
// Returns the persist criterion.
B200_DEV_CALL bool merge_top(const ChainView& c, const SamplerCfg& cfg, int lane) {
  ChainS* s = c.s;
  const int li = s->sp - 2, ri = s->sp - 1;
  double* L = slot(c, cfg, V_STACK + 5 * li);
  double* R = slot(c, cfg, V_STACK + 5 * ri);
  const int ld = cfg.ld;
  const double* im = slot(c, cfg, V_IM);
  const double lsw_l = s->st_lsw[li], lsw_r = s->st_lsw[ri];
  const double lsw_sub = dev_log_sum_exp(lsw_l, lsw_r);
  bool take_right;
  if (lsw_r > lsw_sub) {
    take_right = true;
  } else {
    double accept_prob = dev_exp(lsw_r - lsw_sub);
    take_right = warp_uniform01(s->rng, lane) < accept_prob;
  }
  double a1 = 0, a2 = 0, b1 = 0, b2 = 0, c1 = 0, c2 = 0;
  const bool dense = B200_DENSE(cfg);
  const double *s0 = nullptr, *s1 = nullptr, *s2 = nullptr, *s3 = nullptr;
  if (dense) {  // sharp momenta of the four boundary momenta (the reference keeps them; here they are recomputed)
    dense_sharp(c.mat, L, scratch(c, cfg, 0), cfg.D, lane);
    dense_sharp(c.mat, L + ld, scratch(c, cfg, 1), cfg.D, lane);
    dense_sharp(c.mat, R, scratch(c, cfg, 2), cfg.D, lane);
    dense_sharp(c.mat, R + ld, scratch(c, cfg, 3), cfg.D, lane);
    s0 = scratch(c, cfg, 0); s1 = scratch(c, cfg, 1); s2 = scratch(c, cfg, 2); s3 = scratch(c, cfg, 3);
  }
  B200_VEC_LOOP for (int i = lane; i < cfg.D; i += kWarp) {
    const double m = im[i];
    const double lpb = L[i], lpe = L[ld + i], lr = L[2 * ld + i];
    const double rpb = R[i], rpe = R[ld + i], rr = R[2 * ld + i];
    const double sh_lpb = dense ? s0[i] : m * lpb, sh_lpe = dense ? s1[i] : m * lpe, sh_rpb = dense ? s2[i] : m * rpb,
                 sh_rpe = dense ? s3[i] : m * rpe;
    const double rs = lr + rr;   // rho_subtree = rho_init + rho_final
    a1 += sh_rpe * rs;           // compute_criterion(p_sharp_beg, p_sharp_end, rho_subtree)
    a2 += sh_lpb * rs;
    const double re = lr + rpb;  // rho_init + p_final_beg
    b1 += sh_rpb * re;           // compute_criterion(p_sharp_beg, p_sharp_final_beg, .)
    b2 += sh_lpb * re;
    const double rf = rr + lpe;  // rho_final + p_init_end
    c1 += sh_rpe * rf;           // compute_criterion(p_sharp_init_end, p_sharp_end, .)
    c2 += sh_lpe * rf;
    L[2 * ld + i] = rs;
    L[ld + i] = rpe;
    if (take_right) {
      L[3 * ld + i] = R[3 * ld + i];
      L[4 * ld + i] = R[4 * ld + i];
    }
  }
  B200_SUM6(a1, a2, b1, b2, c1, c2);
  s->st_lsw[li] = lsw_sub;
  if (take_right) {
    s->st_V[li] = s->st_V[ri];
    s->st_H[li] = s->st_H[ri];
    if (B200_DIAG(cfg)) {
      double* pl = diag_slot(c, cfg, 1 + li);
      const double* pr = diag_slot(c, cfg, 1 + ri);
      B200_VEC_LOOP for (int i = lane; i < cfg.D; i += kWarp) pl[i] = pr[i];
    }
  }
  s->sp = ri;
  bool persist = (a1 > 0 && a2 > 0);
  persist &= (b1 > 0 && b2 > 0);
  persist &= (c1 > 0 && c2 > 0);
  return persist;
}

// stepsize_adaptation::learn_stepsize (stepsize_adaptation.hpp:55-71)
B200_DEV_CALL void learn_stepsize(ChainS* s, const SamplerCfg& cfg, double adapt_stat) {
  s->sa_counter += 1.0;
  adapt_stat = adapt_stat > 1 ? 1 : adapt_stat;
  const double eta = 1.0 / (s->sa_counter + cfg.t0);
  s->s_bar = (1.0 - eta) * s->s_bar + eta * (cfg.delta - adapt_stat);
  const double x = s->mu - s->s_bar * sqrt(s->sa_counter) / cfg.gamma;
  const double x_eta = pow(s->sa_counter, -cfg.kappa);
  s->x_bar = (1.0 - x_eta) * s->x_bar + x_eta * x;
  s->eps_nom = dev_exp(x);
}

// var_adaptation::learn_variance (var_adaptation.hpp:17-46) on the new sample q; returns 1 when the metric was
// updated, -1 on numerical overflow of the metric.
B200_DEV_CALL int learn_variance(const ChainView& c, const SamplerCfg& cfg, int lane) {
  ChainS* s = c.s;
  const double* q = slot(c, cfg, V_CQ);
  double* im = slot(c, cfg, V_IM);
  double* wm = slot(c, cfg, V_WM);
  double* wm2 = slot(c, cfg, V_WM2);
  const unsigned cnt = s->win_counter;
  const bool in_window = (cnt >= cfg.w_init_buffer) && (cnt < cfg.w_num_warmup - cfg.w_term_buffer) && (cnt != cfg.w_num_warmup);
  const int D = cfg.D;
  if (in_window && B200_DENSE(cfg)) {  // welford_covar_estimator::add_sample: m2 += (q - m_new) delta^T
    s->wf_n += 1.0;
    const double n = s->wf_n;
    double* delta = scratch(c, cfg, 0);
    double* W = c.mat + 2 * (size_t)D * D;
    B200_VEC_LOOP for (int i = lane; i < D; i += kWarp) {
      const double dl = q[i] - wm[i];
      delta[i] = dl;
      wm[i] += dl / n;
    }
    warp_sync();
    B200_VEC_LOOP for (int i = lane; i < D; i += kWarp) {
      const double r = q[i] - wm[i];
      for (int j = 0; j < D; j++) W[(size_t)j * D + i] += r * delta[j];
    }
  } else if (in_window) {
    s->wf_n += 1.0;
    const double n = s->wf_n;
    B200_VEC_LOOP for (int i = lane; i < cfg.D; i += kWarp) {
      double delta = q[i] - wm[i];
      double mi = wm[i] + delta / n;
      wm[i] = mi;
      wm2[i] += delta * (q[i] - mi);
    }
  }
  const bool end_window = (cnt == s->win_next) && (cnt != cfg.w_num_warmup);
  if (end_window) {
    // compute_next_window (windowed_adaptation.hpp:91-110)
    const unsigned last = cfg.w_num_warmup - cfg.w_term_buffer - 1;
    if (s->win_next != last) {
      s->win_size *= 2;
      s->win_next = cnt + s->win_size;
      if (s->win_next != last) {
        unsigned next_window_boundary = s->win_next + 2 * s->win_size;
        if (next_window_boundary >= cfg.w_num_warmup - cfg.w_term_buffer) s->win_next = last;
      }
    }
    const double n = s->wf_n;
    double bad = 0;
    if (B200_DENSE(cfg)) {  // covar_adaptation::learn_covariance, covar_adaptation.hpp:24-38
      double* M = c.mat;
      double* W = c.mat + 2 * (size_t)D * D;
      B200_VEC_LOOP for (int i = lane; i < D; i += kWarp) {
        for (int j = 0; j < D; j++) {
          const size_t e = (size_t)j * D + i;
          double cov = M[e];
          if (n > 1) cov = W[e] / (n - 1.0);
          cov = (n / (n + 5.0)) * cov + 1e-3 * (5.0 / (n + 5.0)) * (i == j ? 1.0 : 0.0);
          M[e] = cov;
          if (!isfinite(cov)) bad = 1;
          W[e] = 0;
        }
        wm[i] = 0;
      }
      bad = warp_sum(bad);
      s->wf_n = 0;
      s->win_counter = cnt + 1;
      if (bad > 0) return -1;
      return dense_factor(c.mat, D, lane) ? 1 : -1;
    }
    B200_VEC_LOOP for (int i = lane; i < cfg.D; i += kWarp) {
      double var = im[i];
      if (n > 1) var = wm2[i] / (n - 1.0);
      var = (n / (n + 5.0)) * var + 1e-3 * (5.0 / (n + 5.0)) * 1.0;
      im[i] = var;
      if (!isfinite(var)) bad = 1;
      wm[i] = 0;
      wm2[i] = 0;
    }
    bad = warp_sum(bad);
    s->wf_n = 0;
    s->win_counter = cnt + 1;
    return bad > 0 ? -1 : 1;
  }
  s->win_counter = cnt + 1;
  return 0;
}

// util::initialize with a user context (initialize.hpp:93-101), out of line: it runs once per chain and would only bloat the
// state machine.  The random_var_context is constructed (and draws its D uniforms) on every try, also when every value then
// comes from the user's context.  A NaN in the user's vector stands for "not given": the reference then keeps the random
// context's CONSTRAINED value and transform_inits maps it back, so the value is the round trip of the uniform draw through the
// parameter's transform (payload code: 0 none, 1 lower bound 0, 2 unit interval).
B200_DEV_CALL void init_from_user(const ChainView& c, const SamplerCfg& cfg, int lane) {
  ChainS* s = c.s;
  const int D = cfg.D;
  if (cfg.init_radius != 0.0) warp_fill(s->rng, lane, c.q, D, 1, nullptr, cfg.init_radius);
  double partial = 0.0;
  B200_VEC_LOOP for (int i = lane; i < D; i += kWarp) {
    const double user = c.user_q0[i];
    if (user == user) {
      c.q[i] = user;
    } else {
      const double u = cfg.init_radius != 0.0 ? c.q[i] : 0.0;
      const int code = init_nan_code(user);
      c.q[i] = code == 1 ? dev_log(dev_exp(u)) : (code == 2 ? init_logit(dev_inv_logit(u)) : u);
      partial = 1.0;
    }
  }
  s->init_partial = warp_sum(partial) > 0.0 ? 1 : 0;
}

// Resume the chain: consumes the gradient stage's result for z.q when the chain was waiting for one, then runs
// until the next gradient is wanted (ADV_NEED_GRAD) or the chain has nothing left to do (ADV_IDLE).
B200_DEV int advance(const ChainView& c, const SamplerCfg& cfg, int lane) {
  ChainS* s = c.s;
  const int D = cfg.D;
  const double kLog08 = -0x1.c8ff7c79a9a20p-3;  // std::log(0.8), correctly rounded
  for (;;) {
    switch (s->state) {
      case ST_START: {
        s->init_tries = 0;
        s->state = ST_INIT_DRAW;
        if (B200_DENSE(cfg) && !dense_factor(c.mat, D, lane)) {  // "Inverse Euclidean metric not positive definite."
          s->err = 78;
          s->state = ST_DONE;
        }
        break;
      }
      case ST_INIT_DRAW: {  // util::initialize (initialize.hpp:87-117)
        if (c.user_q0) {
          init_from_user(c, cfg, lane);
        } else if (cfg.init_radius == 0.0) {
          B200_VEC_LOOP for (int i = lane; i < D; i += kWarp) c.q[i] = 0.0;
        } else {
          warp_fill(s->rng, lane, c.q, D, 1, nullptr, cfg.init_radius);
        }
        s->init_tries++;
        s->state = ST_INIT_CHECK;
        return ADV_NEED_GRAD;
      }
      case ST_INIT_CHECK: {  // finite lp and finite sum of the gradient (initialize.hpp:148-184)
        double gs = 0;
        B200_VEC_LOOP for (int i = lane; i < D; i += kWarp) gs += c.g[i];
        gs = warp_sum(gs);
        if (isfinite((*c.zV)) && isfinite(gs)) {
          double* cq = slot(c, cfg, V_CQ);
          double* cg = slot(c, cfg, V_CG);
          B200_VEC_LOOP for (int i = lane; i < D; i += kWarp) {
            cq[i] = c.q[i];
            cg[i] = c.g[i];
          }
          s->curV = (*c.zV);
          s->curH = (*c.zV);
          if (cfg.adapt_engaged) {  // run_adaptive_sampler.hpp:60-63
            s->iss_after = 0;
            s->state = ST_ISS_BEGIN;
          } else {
            s->state = ST_ITER_BEGIN;
          }
        } else {
          const int max_tries = ((c.user_q0 && !s->init_partial) || cfg.init_radius == 0.0) ? 1 : 100;
          if (s->init_tries >= max_tries) {
            s->err = 78;  // error_codes::CONFIG -- "Initialization failed."
            s->state = ST_DONE;
          } else {
            s->state = ST_INIT_DRAW;
          }
        }
        break;
      }
      case ST_ISS_BEGIN: {  // base_hmc::init_stepsize, base_hmc.hpp:93-123
        if (s->eps_nom == 0 || s->eps_nom > 1e7 || isnan(s->eps_nom)) {
          s->state = ST_ISS_END;
          break;
        }
        s->iss_H0 = refresh_from_current(c, cfg, lane);
        leapfrog_begin(c, cfg, s->eps_nom, lane);
        s->state = ST_ISS_FIRST;
        return ADV_NEED_GRAD;
      }
      case ST_ISS_FIRST: {
        double hraw;
        double h = leapfrog_end(c, cfg, s->eps_nom, lane, &hraw);
        s->n_grad++;
        double delta_H = s->iss_H0 - h;
        s->iss_dir = delta_H > kLog08 ? 1 : -1;
        s->state = ST_ISS_LOOP_BEGIN;
        break;
      }
      case ST_ISS_LOOP_BEGIN: {  // base_hmc.hpp:125-134
        s->iss_H0 = refresh_from_current(c, cfg, lane);
        leapfrog_begin(c, cfg, s->eps_nom, lane);
        s->state = ST_ISS_LOOP;
        return ADV_NEED_GRAD;
      }
      case ST_ISS_LOOP: {  // base_hmc.hpp:136-158
        double hraw;
        double h = leapfrog_end(c, cfg, s->eps_nom, lane, &hraw);
        s->n_grad++;
        double delta_H = s->iss_H0 - h;
        if ((s->iss_dir == 1) && !(delta_H > kLog08)) {
          s->state = ST_ISS_END;
        } else if ((s->iss_dir == -1) && !(delta_H < kLog08)) {
          s->state = ST_ISS_END;
        } else {
          s->eps_nom = s->iss_dir == 1 ? 2.0 * s->eps_nom : 0.5 * s->eps_nom;
          if (s->eps_nom > 1e7 || s->eps_nom == 0) {
            s->err = 70;  // error_codes::SOFTWARE -- "Posterior is improper" / "No acceptably small step size"
            s->state = ST_DONE;
          } else {
            s->state = ST_ISS_LOOP_BEGIN;
          }
        }
        break;
      }
      case ST_ISS_END: {  // z restored to z_init (the current sample is untouched); adapt_diag_e_nuts.hpp:35-40
        if (s->iss_after == 1) {
          if (B200_STATIC(cfg)) {  // ... and after the init_stepsize that follows a metric update
            const int L = (int)(cfg.int_time / s->eps_nom);
            s->st_L = L < 1 ? 1 : L;
          }
          s->mu = dev_log(10 * s->eps_nom);
          s->sa_counter = 0;
          s->s_bar = 0;
          s->x_bar = 0;
          s->state = ST_ITER_NEXT;
        } else {
          s->state = ST_ITER_BEGIN;
        }
        break;
      }
      case ST_ITER_BEGIN: {  // base_nuts::transition, base_nuts.hpp:78-121
        if (s->iter >= s->iter_end) return ADV_IDLE;
        s->eps = s->eps_nom;  // sample_stepsize, base_hmc.hpp:195-200
        if (cfg.stepsize_jitter != 0.0) s->eps *= 1.0 + cfg.stepsize_jitter * (2.0 * warp_uniform01(s->rng, lane) - 1.0);
        s->H0 = refresh_from_current(c, cfg, lane);
        if (B200_STATIC(cfg)) {  // base_static_hmc::transition: z_init is the current sample (kept in V_CQ / V_CG / curV)
          s->st_i = 0;
          s->state = ST_STATIC_STEP_BEGIN;
          break;
        }
        {
          const double* p = slot(c, cfg, V_P);
          double* fq = slot(c, cfg, V_FQ); double* fp = slot(c, cfg, V_FP); double* fg = slot(c, cfg, V_FG);
          double* bq = slot(c, cfg, V_BQ); double* bp = slot(c, cfg, V_BP); double* bg = slot(c, cfg, V_BG);
          double* pfb = slot(c, cfg, V_PFB); double* pbf = slot(c, cfg, V_PBF); double* rho = slot(c, cfg, V_RHO);
          B200_VEC_LOOP for (int i = lane; i < D; i += kWarp) {
            const double qi = c.q[i], pi = p[i], gi = c.g[i];
            fq[i] = qi; fp[i] = pi; fg[i] = gi;
            bq[i] = qi; bp[i] = pi; bg[i] = gi;
            pfb[i] = pi; pbf[i] = pi; rho[i] = pi;
          }
          if (B200_DIAG(cfg)) {  // z_sample starts as the initial point with its fresh momentum
            double* cp = diag_slot(c, cfg, 0);
            B200_VEC_LOOP for (int i = lane; i < D; i += kWarp) cp[i] = p[i];
          }
        }
        s->curH = s->H0;  // energy of z_sample while it is still the initial point
        s->lsw = 0;
        s->n_leapfrog = 0;
        s->sum_metro = 0;
        s->depth = 0;
        s->divergent = 0;
        s->state = ST_DOUBLING_BEGIN;
        break;
      }
      case ST_DOUBLING_BEGIN: {  // base_nuts.hpp:123-157
        if (s->depth >= cfg.max_depth) {
          s->state = ST_TRANS_END;
          break;
        }
        const bool fwd = warp_uniform01(s->rng, lane) > 0.5;
        s->dir = fwd ? 1 : -1;
        {
          double* p = slot(c, cfg, V_P);
          const double* eq = slot(c, cfg, fwd ? V_FQ : V_BQ);
          const double* ep = slot(c, cfg, fwd ? V_FP : V_BP);
          const double* eg = slot(c, cfg, fwd ? V_FG : V_BG);
          double* pmid = slot(c, cfg, fwd ? V_PBF : V_PFB);  // p_bck_fwd = p_fwd_fwd  /  p_fwd_bck = p_bck_bck
          B200_VEC_LOOP for (int i = lane; i < D; i += kWarp) {
            c.q[i] = eq[i];
            p[i] = ep[i];
            c.g[i] = eg[i];
            pmid[i] = ep[i];
          }
        }
        s->leaf_count = 0;
        s->n_leaves = 1u << s->depth;
        s->sp = 0;
        s->state = ST_LEAF_BEGIN;
        break;
      }
      case ST_LEAF_BEGIN: {
        leapfrog_begin(c, cfg, s->dir * s->eps, lane);
        s->state = ST_LEAF;
        return ADV_NEED_GRAD;
      }
      case ST_LEAF: {  // build_tree base case, base_nuts.hpp:253-282, followed by the pending merges
        double hraw;
        const double h = leapfrog_end(c, cfg, s->dir * s->eps, lane, &hraw);
        s->n_grad++;
        s->n_leapfrog++;
        if ((h - s->H0) > 1000.0) s->divergent = 1;
        const double w = s->H0 - h;
        if (w > 0) s->sum_metro += 1;
        else s->sum_metro += dev_exp(w);
        bool valid = !s->divergent;
        {
          const int e = s->sp;
          double* E = slot(c, cfg, V_STACK + 5 * e);
          const double* p = slot(c, cfg, V_P);
          const int ld = cfg.ld;
          B200_VEC_LOOP for (int i = lane; i < D; i += kWarp) {
            const double pi = p[i];
            E[i] = pi;
            E[ld + i] = pi;
            E[2 * ld + i] = pi;
            E[3 * ld + i] = c.q[i];
            E[4 * ld + i] = c.g[i];
          }
          if (B200_DIAG(cfg)) {
            double* pp = diag_slot(c, cfg, 1 + e);
            B200_VEC_LOOP for (int i = lane; i < D; i += kWarp) pp[i] = p[i];
          }
          s->st_lsw[e] = w;  // log_sum_exp(-inf, H0 - h)
          s->st_V[e] = (*c.zV);
          s->st_H[e] = hraw;
          s->sp = e + 1;
        }
        s->leaf_count++;
        if (valid) {
          unsigned lc = s->leaf_count;
          while (valid && (lc & 1u) == 0u) {
            valid = merge_top(c, cfg, lane);
            lc >>= 1;
          }
        }
        if (valid && s->leaf_count < s->n_leaves) {
          s->state = ST_LEAF_BEGIN;
          break;
        }
        s->sub_valid = valid ? 1 : 0;
        s->state = ST_SUBTREE_END;
        break;
      }
      case ST_SUBTREE_END: {  // subtree finished (or aborted): base_nuts.hpp:142,154 then :156-192
        const bool valid = s->sub_valid != 0;
        const bool fwd = s->dir == 1;
        {
          const double* p = slot(c, cfg, V_P);
          double* eq = slot(c, cfg, fwd ? V_FQ : V_BQ);
          double* ep = slot(c, cfg, fwd ? V_FP : V_BP);
          double* eg = slot(c, cfg, fwd ? V_FG : V_BG);
          B200_VEC_LOOP for (int i = lane; i < D; i += kWarp) {
            eq[i] = c.q[i];
            ep[i] = p[i];
            eg[i] = c.g[i];
          }
        }
        if (!valid) {
          s->state = ST_TRANS_END;
          break;
        }
        s->depth++;
        const double lsw_sub = s->st_lsw[0];
        bool take;
        if (lsw_sub > s->lsw) {
          take = true;
        } else {
          double accept_prob = dev_exp(lsw_sub - s->lsw);
          take = warp_uniform01(s->rng, lane) < accept_prob;
        }
        s->lsw = dev_log_sum_exp(s->lsw, lsw_sub);
        double a1 = 0, a2 = 0, b1 = 0, b2 = 0, c1 = 0, c2 = 0;
        {
          const int ld = cfg.ld;
          double* S = slot(c, cfg, V_STACK);  // the finished subtree: p_beg, p_end, rho, proposal q, g
          const double* im = slot(c, cfg, V_IM);
          double* rho = slot(c, cfg, V_RHO);
          double* pfb = slot(c, cfg, V_PFB);
          double* pbf = slot(c, cfg, V_PBF);
          const double* pff = slot(c, cfg, V_FP);  // p_fwd_fwd == z_fwd.p
          const double* pbb = slot(c, cfg, V_BP);  // p_bck_bck == z_bck.p
          double* cq = slot(c, cfg, V_CQ);
          double* cg = slot(c, cfg, V_CG);
          const bool dense = B200_DENSE(cfg);
          const double *s0 = nullptr, *s1 = nullptr, *s2 = nullptr, *s3 = nullptr;
          if (dense) {  // boundary momentum of the new subtree first, then the four sharp momenta
            double* dst = fwd ? pfb : pbf;
            B200_VEC_LOOP for (int i = lane; i < D; i += kWarp) dst[i] = S[i];
            dense_sharp(c.mat, pff, scratch(c, cfg, 0), D, lane);
            dense_sharp(c.mat, pbb, scratch(c, cfg, 1), D, lane);
            dense_sharp(c.mat, pfb, scratch(c, cfg, 2), D, lane);
            dense_sharp(c.mat, pbf, scratch(c, cfg, 3), D, lane);
            s0 = scratch(c, cfg, 0); s1 = scratch(c, cfg, 1); s2 = scratch(c, cfg, 2); s3 = scratch(c, cfg, 3);
          }
          B200_VEC_LOOP for (int i = lane; i < D; i += kWarp) {
            const double m = im[i];
            double rho_bck, rho_fwd;
            if (fwd) {
              rho_bck = rho[i];
              rho_fwd = S[2 * ld + i];
              if (!dense) pfb[i] = S[i];  // dense: stored above, and other lanes may still be reading it in dense_sharp
            } else {
              rho_fwd = rho[i];
              rho_bck = S[2 * ld + i];
              if (!dense) pbf[i] = S[i];
            }
            const double v_pfb = pfb[i], v_pbf = pbf[i], v_pff = pff[i], v_pbb = pbb[i];
            const double sh_pff = dense ? s0[i] : m * v_pff, sh_pbb = dense ? s1[i] : m * v_pbb, sh_pfb = dense ? s2[i] : m * v_pfb,
                         sh_pbf = dense ? s3[i] : m * v_pbf;
            const double r = rho_bck + rho_fwd;
            a1 += sh_pff * r;  // compute_criterion(p_sharp_bck_bck, p_sharp_fwd_fwd, rho)
            a2 += sh_pbb * r;
            const double re = rho_bck + v_pfb;
            b1 += sh_pfb * re; // compute_criterion(p_sharp_bck_bck, p_sharp_fwd_bck, rho_bck + p_fwd_bck)
            b2 += sh_pbb * re;
            const double rf = rho_fwd + v_pbf;
            c1 += sh_pff * rf; // compute_criterion(p_sharp_bck_fwd, p_sharp_fwd_fwd, rho_fwd + p_bck_fwd)
            c2 += sh_pbf * rf;
            rho[i] = r;
            if (take) {
              cq[i] = S[3 * ld + i];
              cg[i] = S[4 * ld + i];
            }
          }
        }
        if (take) {
          s->curV = s->st_V[0];
          s->curH = s->st_H[0];
          if (B200_DIAG(cfg)) {
            double* cp = diag_slot(c, cfg, 0);
            const double* pp = diag_slot(c, cfg, 1);
            B200_VEC_LOOP for (int i = lane; i < D; i += kWarp) cp[i] = pp[i];
          }
        }
        B200_SUM6(a1, a2, b1, b2, c1, c2);
        bool persist = (a1 > 0 && a2 > 0);
        persist &= (b1 > 0 && b2 > 0);
        persist &= (c1 > 0 && c2 > 0);
        s->state = persist ? ST_DOUBLING_BEGIN : ST_TRANS_END;
        break;
      }
#if B200_WITH_DENSE  // static HMC lives only in the full-featured kernels (dense.cu)
      case ST_STATIC_STEP_BEGIN: {
        if (s->st_i >= s->st_L) {
          s->state = ST_STATIC_END;
          break;
        }
        leapfrog_begin(c, cfg, s->eps, lane);
        s->state = ST_STATIC_STEP;
        return ADV_NEED_GRAD;
      }
      case ST_STATIC_STEP: {
        double hraw;
        s->lsw = leapfrog_end(c, cfg, s->eps, lane, &hraw);  // H of the working point (NaN -> +inf); raw value for energy__
        s->curH = hraw;
        s->n_grad++;
        s->st_i++;
        s->state = ST_STATIC_STEP_BEGIN;
        break;
      }
      case ST_STATIC_END: {  // base_static_hmc.hpp:56-69
        const double h = s->lsw;
        double accept_prob = dev_exp(s->H0 - h);
        bool reject = false;
        if (accept_prob < 1) reject = warp_uniform01(s->rng, lane) > accept_prob;
        if (reject) {
          s->curH = s->H0;  // H(z_init) with the momentum drawn for this transition
        } else {
          double* cq = slot(c, cfg, V_CQ);
          double* cg = slot(c, cfg, V_CG);
          B200_VEC_LOOP for (int i = lane; i < D; i += kWarp) {
            cq[i] = c.q[i];
            cg[i] = c.g[i];
          }
          s->curV = (*c.zV);
        }
        s->sum_metro = accept_prob > 1 ? 1 : accept_prob;
        s->n_leapfrog = 1;      // accept_stat = sum_metro / n_leapfrog in ST_TRANS_END
        s->depth = s->st_L;     // row column 3: L (the writer prints int_time__ instead)
        s->divergent = 0;
        s->state = ST_TRANS_END;
        break;
      }
#endif
      case ST_TRANS_END: {  // base_nuts.hpp:194-203, generate_transitions.hpp:70-75, adapt_diag_e_nuts.hpp:28-41
        const double accept_stat = s->sum_metro / (double)s->n_leapfrog;
        const bool warm = s->iter < cfg.num_warmup;
        const int m = warm ? s->iter : s->iter - cfg.num_warmup;
        const bool save = (warm ? cfg.save_warmup != 0 : true) && (m % cfg.num_thin == 0);
        if (save && c.draws && s->draws_row < cfg.rows_per_chain) {
          double* r = c.draws + (size_t)s->draws_row * cfg.ncol;
          if (lane == 0) {
            r[0] = -s->curV;
            r[1] = accept_stat;
            r[2] = s->eps;
            r[3] = s->depth;
            r[4] = s->n_leapfrog;
            r[5] = s->divergent;
            r[6] = s->curH;
          }
          const double* cq = slot(c, cfg, V_CQ);
          B200_VEC_LOOP for (int i = lane; i < D; i += kWarp) r[7 + i] = cq[i];
          if (B200_DIAG(cfg)) {
            const double* cp = diag_slot(c, cfg, 0);
            const double* cg = slot(c, cfg, V_CG);
            B200_VEC_LOOP for (int i = lane; i < D; i += kWarp) {
              r[7 + D + i] = cp[i];
              r[7 + 2 * D + i] = cg[i];
            }
          }
          s->draws_row++;
        }
        s->iter++;
        if (cfg.adapt_engaged && warm) {
          learn_stepsize(s, cfg, accept_stat);
          if (B200_STATIC(cfg)) {  // adapt_*_static_hmc::transition: update_L_() after every step-size update
            const int L = (int)(cfg.int_time / s->eps_nom);
            s->st_L = L < 1 ? 1 : L;
          }
          const int upd = learn_variance(c, cfg, lane);
          if (upd < 0) {
            s->err = 70;  // "Numerical overflow in metric adaptation."
            s->state = ST_DONE;
            break;
          }
          if (upd > 0) {
            s->iss_after = 1;
            s->state = ST_ISS_BEGIN;
            break;
          }
        }
        s->state = ST_ITER_NEXT;
        break;
      }
      case ST_ITER_NEXT: {  // disengage_adaptation after the last warm-up transition (adapt_diag_e_nuts.hpp:45-48)
        if (cfg.adapt_engaged && s->iter == cfg.num_warmup) s->eps_nom = dev_exp(s->x_bar);
        s->state = ST_ITER_BEGIN;
        break;
      }
      default:
        return ADV_IDLE;
    }
  }
}

// set-up of the sampler scalars, hmc_nuts_diag_e_adapt.hpp:84-98 + base_hmc ctor + stepsize_adaptation ctor
B200_HD void chain_init_scalars(ChainS* s, const SamplerCfg& cfg) {
  s->state = ST_START;
  s->err = 0;
  s->iter = 0;
  s->iter_end = 0;
  s->init_tries = 0;
  s->init_partial = 0;
  s->st_i = 0;
  s->sub_valid = 0;
  {  // base_static_hmc::set_nominal_stepsize_and_T(stepsize, int_time): L from the stepsize ARGUMENT
    int L = cfg.static_hmc ? (int)(cfg.int_time / (cfg.stepsize > 0 ? cfg.stepsize : 0.1)) : 0;
    s->st_L = L < 1 ? 1 : L;
  }
  s->iss_after = 0;
  s->iss_dir = 0;
  s->dir = 0;
  s->depth = 0;
  s->n_leapfrog = 0;
  s->divergent = 0;
  s->sp = 0;
  s->leaf_count = 0;
  s->n_leaves = 0;
  s->draws_row = 0;
  s->n_grad = 0;
  s->eps_nom = cfg.stepsize > 0 ? cfg.stepsize : 0.1;
  s->eps = s->eps_nom;
  s->curV = 0; s->curH = 0; s->H0 = 0; s->lsw = 0; s->sum_metro = 0; s->iss_H0 = 0;
  s->sa_counter = 0; s->s_bar = 0; s->x_bar = 0;
  s->mu = log(10 * cfg.stepsize);
  s->wf_n = 0;
  s->win_counter = 0;
  s->win_size = cfg.w_size0;
  s->win_next = cfg.w_next0;
}

}  // namespace b200
