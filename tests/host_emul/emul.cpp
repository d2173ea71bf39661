// TEST-ONLY: builds the device state machine (cmdstan_b200/csrc/nuts.cuh) for the host with a one-lane "warp"
// (B200_HOST_EMUL) so that the CPU suite can check its logic against the oracle without a GPU.
// This file is never part of libb200nuts.so; the product has no CPU path.
#include <cmath>
#include <cstring>
#include <vector>
using std::isnan; using std::isfinite;
#include "../../cmdstan_b200/csrc/nuts.cuh"
#include "../../cmdstan_b200/csrc/host_seed.h"

using namespace b200;

extern "C" int emul_run_chain(int kind, int D, int N, const double* yr, const double* sigma, int sum_y,
                              unsigned seed, unsigned chain_id, int num_warmup, int num_samples, int num_thin, int save_warmup,
                              double stepsize, double jitter, int max_depth, double delta, double gamma, double kappa, double t0,
                              unsigned init_buffer, unsigned term_buffer, unsigned window, double init_radius, int adapt_engaged,
                              const double* q0, const double* inv_metric0, double* draws, double* adapt_out, long long* n_grad,
                              int unit_e, int dense_e, int diag, int static_hmc, double int_time) {
  ModelData m{};
  m.kind = kind; m.D = D; m.N = N; m.yr = yr; m.sigma = sigma; m.sum_y = sum_y;
  SamplerCfg cfg{};
  cfg.D = D; cfg.ld = D; cfg.max_depth = max_depth;
  cfg.num_warmup = num_warmup; cfg.num_samples = num_samples; cfg.num_thin = num_thin; cfg.save_warmup = save_warmup;
  cfg.adapt_engaged = adapt_engaged; cfg.user_inits = q0 != nullptr;
  cfg.stepsize = stepsize; cfg.stepsize_jitter = (jitter > 0 && jitter < 1) ? jitter : 0.0;
  cfg.delta = (delta > 0 && delta < 1) ? delta : 0.5; cfg.gamma = gamma > 0 ? gamma : 0.05;
  cfg.kappa = kappa > 0 ? kappa : 0.75; cfg.t0 = t0 > 0 ? t0 : 10; cfg.init_radius = init_radius;
  cfg.dense = dense_e;
  cfg.diag = diag;
  cfg.static_hmc = static_hmc;
  cfg.int_time = int_time;
  resolve_window(&cfg, (adapt_engaged && !unit_e) ? num_warmup : 0, init_buffer, term_buffer, window);
  cfg.ncol = 7 + D * (diag ? 3 : 1);
  cfg.rows_per_chain = (save_warmup ? (num_warmup + num_thin - 1) / num_thin : 0) + (num_samples + num_thin - 1) / num_thin;
  std::vector<double> slab((size_t)num_vec_slots(max_depth, dense_e, diag) * D, 0.0), q(D), g(D), mat(dense_e ? 3 * (size_t)D * D : 1, 0.0);
  ChainS s;
  std::memset(&s, 0, sizeof s);
  chain_init_scalars(&s, cfg);
  seed_stream(s.rng, seed, chain_id);
  double zV = 0;
  ChainView c{&s, q.data(), g.data(), &zV, slab.data(), draws, q0};
  if (dense_e) {  // inv_metric0: [D][D] row-major -> column-major
    c.mat = mat.data();
    for (int i = 0; i < D; i++)
      for (int j = 0; j < D; j++) mat[(size_t)j * D + i] = inv_metric0 ? inv_metric0[(size_t)i * D + j] : (i == j ? 1.0 : 0.0);
  } else {
    for (int i = 0; i < D; i++) slab[(size_t)V_IM * D + i] = inv_metric0 ? inv_metric0[i] : 1.0;
  }
  s.iter_end = num_warmup + num_samples;
  while (advance(c, cfg, 0) == ADV_NEED_GRAD) {
    double lp = lp_grad_small(m, q.data(), g.data(), 0);
    for (int i = 0; i < D; i++) g[i] = -g[i];
    zV = -lp;
  }
  adapt_out[0] = s.eps_nom;
  if (dense_e) {
    for (int i = 0; i < D; i++)
      for (int j = 0; j < D; j++) adapt_out[1 + (size_t)i * D + j] = mat[(size_t)j * D + i];
  } else {
    for (int i = 0; i < D; i++) adapt_out[1 + i] = slab[(size_t)V_IM * D + i];
  }
  *n_grad = s.n_grad;
  return s.err;
}

// state-injection KAT entry points (same cases as the oracle's, against the device code)
extern "C" void emul_learn_stepsize_kat(double counter, double s_bar, double x_bar, double mu, double delta, double gamma,
                                        double kappa, double t0, double adapt_stat, double* out) {
  ChainS s;
  std::memset(&s, 0, sizeof s);
  SamplerCfg cfg{};
  s.sa_counter = counter; s.s_bar = s_bar; s.x_bar = x_bar; s.mu = mu;
  cfg.delta = delta; cfg.gamma = gamma; cfg.kappa = kappa; cfg.t0 = t0;
  learn_stepsize(&s, cfg, adapt_stat);
  out[0] = s.eps_nom; out[1] = s.sa_counter; out[2] = s.s_bar; out[3] = s.x_bar;
}

extern "C" int emul_learn_variance_kat(int D, const double* q, int n_learn, unsigned num_warmup, unsigned init_buffer,
                                       unsigned term_buffer, unsigned window, double* var_out) {
  SamplerCfg cfg{};
  cfg.D = D; cfg.ld = D;
  resolve_window(&cfg, num_warmup, init_buffer, term_buffer, window);
  std::vector<double> slab((size_t)num_vec_slots(1) * D, 0.0), qq(D), g(D);
  ChainS s;
  std::memset(&s, 0, sizeof s);
  s.win_size = cfg.w_size0; s.win_next = cfg.w_next0;
  double zV = 0;
  ChainView c{&s, qq.data(), g.data(), &zV, slab.data(), nullptr, nullptr};
  for (int i = 0; i < D; i++) slab[(size_t)V_CQ * D + i] = q[i];
  int updates = 0;
  for (int i = 0; i < n_learn; i++) updates += learn_variance(c, cfg, 0) > 0;
  for (int i = 0; i < D; i++) var_out[i] = slab[(size_t)V_IM * D + i];
  return updates;
}

// covar_adaptation_test.cpp restated on the device code (dense windows: Welford covariance in the chain's matrix block)
extern "C" int emul_learn_covariance_kat(int D, const double* q, int n_learn, unsigned num_warmup, unsigned init_buffer,
                                         unsigned term_buffer, unsigned window, double* covar_out) {
  SamplerCfg cfg{};
  cfg.D = D; cfg.ld = D; cfg.dense = 1;
  resolve_window(&cfg, num_warmup, init_buffer, term_buffer, window);
  std::vector<double> slab((size_t)num_vec_slots(1, 1, 0) * D, 0.0), qq(D), g(D), mat(3 * (size_t)D * D, 0.0);
  ChainS s;
  std::memset(&s, 0, sizeof s);
  s.win_size = cfg.w_size0; s.win_next = cfg.w_next0;
  double zV = 0;
  ChainView c{&s, qq.data(), g.data(), &zV, slab.data(), nullptr, nullptr};
  c.mat = mat.data();
  for (int i = 0; i < D; i++) slab[(size_t)V_CQ * D + i] = q[i];
  int updates = 0;
  for (int i = 0; i < n_learn; i++) updates += learn_variance(c, cfg, 0) > 0;
  for (int i = 0; i < D; i++)
    for (int j = 0; j < D; j++) covar_out[(size_t)i * D + j] = mat[(size_t)j * D + i];
  return updates;
}

// expl_leapfrog::evolve through the DEVICE code path (leapfrog_begin, model gradient, leapfrog_end) for the std-normal model:
// same KATs as the oracle's oracle_evolve_kat.  out = q[D], p[D], g[D], V, H
extern "C" void emul_evolve_kat(int D, const double* q, const double* p, const double* inv_metric, double eps, double* out) {
  ModelData m{};
  m.kind = MODEL_STD_NORMAL; m.D = D;
  SamplerCfg cfg{};
  cfg.D = D; cfg.ld = D; cfg.max_depth = 1;
  std::vector<double> slab((size_t)num_vec_slots(1) * D, 0.0), qq(q, q + D), g(D);
  ChainS s;
  std::memset(&s, 0, sizeof s);
  double zV = 0;
  ChainView c{&s, qq.data(), g.data(), &zV, slab.data(), nullptr, nullptr};
  for (int i = 0; i < D; i++) {
    slab[(size_t)V_IM * D + i] = inv_metric ? inv_metric[i] : 1.0;
    slab[(size_t)V_P * D + i] = p[i];
  }
  auto grad = [&]() {
    double lp = lp_grad_small(m, qq.data(), g.data(), 0);
    for (int i = 0; i < D; i++) g[i] = -g[i];
    zV = -lp;
  };
  grad();
  leapfrog_begin(c, cfg, eps, 0);
  grad();
  double hraw;
  const double H = leapfrog_end(c, cfg, eps, 0, &hraw);
  for (int i = 0; i < D; i++) {
    out[i] = qq[i];
    out[D + i] = slab[(size_t)V_P * D + i];
    out[2 * D + i] = g[i];
  }
  out[3 * D] = zV;
  out[3 * D + 1] = H;
}
