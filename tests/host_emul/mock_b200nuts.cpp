// TEST-ONLY mock of the C ABI (include/b200nuts.h) on top of the CPU oracle (oracle/liboracle.so), built as
// tests/host_emul/mock/libb200nuts.so.  It exists so that the C++ host side above the ABI -- cmdstan_b200/host (argument tree,
// service shim, Stan-CSV writer, metric JSON, file naming, thinning, progress messages) -- can be run end to end on a machine
// without a GPU and compared byte for byte with the UNMODIFIED reference (tests/test_cli_mock_vs_reference.py).  It is never
// built into, linked with or loaded by the product; cmdstan_b200/libb200nuts.so has no CPU path.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/b200nuts.h"
#include "../../include/b200nuts_model.h"
#include "../../oracle/nuts_oracle.h"

struct b200nuts_engine {
  b200nuts_config cfg;
  std::string err;
  model_t m{};
  std::vector<double> X, yr, sigma;
  std::vector<int> yi;
  std::vector<signed char> y8;
  int std_normal_D = 3;
  bool ready = false;
  std::vector<double> q0, im;
  // per chain: all saved rows [rows][7 + D], adaptation [1 + D or D*D], rc
  std::vector<std::vector<double>> draws, adapt;
  std::vector<int> rc;
  std::vector<long> ngrad;
  int64_t rows_per_chain = 0;
  int iters_done = 0;
};

static std::string g_err;

static int finalize(b200nuts_engine* e) {
  model_t& m = e->m;
  m.kind = e->cfg.model;
  switch (e->cfg.model) {
    case B200NUTS_STD_NORMAL: m.D = e->std_normal_D; break;
    case B200NUTS_BERNOULLI:
      if (e->yi.empty()) { e->err = "bernoulli: data variable y missing"; return B200NUTS_NOINPUT; }
      m.D = 1; m.N = (int)e->yi.size(); m.yi = e->yi.data(); break;
    case B200NUTS_EIGHT_SCHOOLS_NC:
      if (e->yr.empty() || e->yr.size() != e->sigma.size()) { e->err = "eight_schools: y and sigma"; return B200NUTS_NOINPUT; }
      m.N = (int)e->yr.size(); m.D = m.N + 2; m.yr = e->yr.data(); m.sigma = e->sigma.data(); break;
    case B200NUTS_LOGISTIC_GLM:
    case B200NUTS_POISSON_GLM:
      m.D = m.K + 1; m.X = e->X.data(); m.yi = e->yi.data(); break;
    case B200NUTS_IRT_2PL:
      m.D = 4 + m.N + 2 * m.K; m.y8 = e->y8.data(); break;
    default: e->err = "mock: unknown model"; return B200NUTS_USAGE;
  }
  e->ready = true;
  return 0;
}

extern "C" {

void b200nuts_default_config(b200nuts_config* c) {
  std::memset(c, 0, sizeof *c);
  c->num_chains = 1; c->init_chain_id = 1; c->init_radius = 2.0; c->num_warmup = 1000; c->num_samples = 1000; c->num_thin = 1;
  c->stepsize = 1.0; c->max_depth = 10; c->delta = 0.8; c->gamma = 0.05; c->kappa = 0.75; c->t0 = 10.0;
  c->init_buffer = 75; c->term_buffer = 50; c->window = 25; c->adapt_engaged = 1; c->store_draws = 1; c->row_shards = 1;
  c->int_time = 6.28318530717958647692;
}
const char* b200nuts_last_error(const b200nuts_engine* e) { return e ? e->err.c_str() : g_err.c_str(); }
int b200nuts_create(const b200nuts_config* cfg, b200nuts_engine** out) {
  *out = nullptr;
  if (cfg->num_chains < 1 || cfg->model < 0 || cfg->model > 5) { g_err = "mock: bad config"; return B200NUTS_USAGE; }
  if (cfg->adapt_engaged && cfg->num_warmup == 0) { g_err = "Model is adapting but num_warmup is 0"; return B200NUTS_USAGE; }
  auto* e = new b200nuts_engine();
  e->cfg = *cfg;
  *out = e;
  return 0;
}
void b200nuts_destroy(b200nuts_engine* e) { delete e; }
static int64_t prod(const int64_t* d, int nd) { int64_t p = 1; for (int i = 0; i < nd; i++) p *= d[i]; return p; }
int b200nuts_set_data_f64(b200nuts_engine* e, const char* name, const double* h, const int64_t* dims, int nd) {
  const std::string n(name);
  const int64_t cnt = prod(dims, nd);
  if (n == "X" && nd == 2) { e->X.assign(h, h + cnt); e->m.N = (int)dims[0]; e->m.K = (int)dims[1]; }
  else if (n == "y") e->yr.assign(h, h + cnt);
  else if (n == "sigma") e->sigma.assign(h, h + cnt);
  else { e->err = "mock: unknown f64 variable " + n; return B200NUTS_DATAERR; }
  e->ready = false;
  return 0;
}
int b200nuts_set_data_i32(b200nuts_engine* e, const char* name, const int32_t* h, const int64_t* dims, int nd) {
  const std::string n(name);
  const int64_t cnt = prod(dims, nd);
  if (n == "y") e->yi.assign(h, h + cnt);
  else if (n == "D" && cnt == 1) e->std_normal_D = h[0];
  else { e->err = "mock: unknown i32 variable " + n; return B200NUTS_DATAERR; }
  e->ready = false;
  return 0;
}
int b200nuts_set_data_i8(b200nuts_engine* e, const char* name, const int8_t* h, const int64_t* dims, int nd) {
  if (std::string(name) != "y" || nd != 2) { e->err = "mock: unknown i8 variable"; return B200NUTS_DATAERR; }
  e->y8.assign(h, h + prod(dims, nd)); e->m.N = (int)dims[0]; e->m.K = (int)dims[1];
  e->ready = false;
  return 0;
}
int b200nuts_num_params(const b200nuts_engine* ce, int64_t* nu, int64_t* no) {
  auto* e = const_cast<b200nuts_engine*>(ce);
  if (!e->ready) { int rc = finalize(e); if (rc) return rc; }
  if (nu) *nu = e->m.D;
  if (no) *no = oracle_num_outputs(&e->m);
  return 0;
}
int b200nuts_set_inits(b200nuts_engine* e, const double* q0) {
  if (!e->ready) { int rc = finalize(e); if (rc) return rc; }
  if (q0) e->q0.assign(q0, q0 + (size_t)e->cfg.num_chains * e->m.D); else e->q0.clear();
  return 0;
}
int b200nuts_set_inv_metric(b200nuts_engine* e, const double* im) {
  if (!e->ready) { int rc = finalize(e); if (rc) return rc; }
  const size_t md = e->cfg.metric == 2 ? (size_t)e->m.D * e->m.D : (size_t)e->m.D;
  if (im) e->im.assign(im, im + (size_t)e->cfg.num_chains * md); else e->im.clear();
  return 0;
}
int b200nuts_init_chains(b200nuts_engine* e) {
  if (!e->ready) { int rc = finalize(e); if (rc) return rc; }
  const b200nuts_config& c = e->cfg;
  const int D = e->m.D, C = c.num_chains;
  const size_t md = c.metric == 2 ? (size_t)D * D : (size_t)D;
  e->rows_per_chain = (c.save_warmup ? (c.num_warmup + c.num_thin - 1) / c.num_thin : 0) + (c.num_samples + c.num_thin - 1) / c.num_thin;
  const int ncol = 7 + D * (c.store_diagnostics ? 3 : 1);
  e->draws.assign(C, std::vector<double>((size_t)e->rows_per_chain * ncol + 1, 0.0));
  e->adapt.assign(C, std::vector<double>(1 + md, 0.0));
  e->rc.assign(C, 0);
  e->ngrad.assign(C, 0);
  for (int k = 0; k < C; k++) {
    oracle_config_t oc;
    std::memset(&oc, 0, sizeof oc);
    oc.seed = c.random_seed; oc.chain_id = c.init_chain_id + (uint32_t)k;
    oc.num_warmup = c.num_warmup; oc.num_samples = c.num_samples; oc.num_thin = c.num_thin; oc.save_warmup = c.save_warmup;
    oc.stepsize = c.stepsize; oc.stepsize_jitter = c.stepsize_jitter; oc.max_depth = c.max_depth;
    oc.delta = c.delta; oc.gamma = c.gamma; oc.kappa = c.kappa; oc.t0 = c.t0;
    oc.init_buffer = c.init_buffer; oc.term_buffer = c.term_buffer; oc.window = c.window;
    oc.init_radius = c.init_radius; oc.adapt_engaged = c.adapt_engaged; oc.unit_e = c.metric == 1; oc.dense_e = c.metric == 2;
    oc.diag = c.store_diagnostics;
    oc.static_hmc = c.static_hmc;
    oc.int_time = c.int_time;
    e->rc[k] = oracle_run_chain(&e->m, &oc, e->q0.empty() ? nullptr : e->q0.data() + (size_t)k * D,
                                e->im.empty() ? nullptr : e->im.data() + (size_t)k * md, e->draws[k].data(), e->adapt[k].data(), nullptr,
                                &e->ngrad[k]);
    if (e->rc[k]) {  // adaptation of a failed chain: whatever it had
      e->adapt[k][0] = c.stepsize;
    }
  }
  e->iters_done = 0;
  return 0;
}
int b200nuts_log_prob_grad(b200nuts_engine* e, const double* q, int64_t n, int jacobian, double* lp, double* grad) {
  if (!e->ready) { int rc = finalize(e); if (rc) return rc; }
  if (!jacobian) { e->err = "mock: jacobian = 0 is not available"; return B200NUTS_USAGE; }
  for (int64_t i = 0; i < n; i++) lp[i] = oracle_lp_grad(&e->m, q + i * e->m.D, grad + i * e->m.D);
  return 0;
}
int b200nuts_run(b200nuts_engine* e, int n_iter) {
  if (e->draws.empty()) { e->err = "b200nuts_init_chains must be called first"; return B200NUTS_USAGE; }
  e->iters_done = std::min(e->iters_done + n_iter, e->cfg.num_warmup + e->cfg.num_samples);
  return 0;
}
int b200nuts_fetch_draws(b200nuts_engine* e, int64_t chain0, int64_t nchains, int64_t row0, int64_t nrows, int constrained, double* out) {
  const int D = e->m.D, K = constrained ? oracle_num_outputs(&e->m) : D;
  if (chain0 < 0 || chain0 + nchains > e->cfg.num_chains || row0 < 0 || row0 + nrows > e->rows_per_chain) { e->err = "fetch_draws: range out of bounds"; return B200NUTS_USAGE; }
  for (int64_t c = 0; c < nchains; c++)
    for (int64_t r = 0; r < nrows; r++) {
      const double* src = e->draws[chain0 + c].data() + (size_t)(row0 + r) * (7 + D * (e->cfg.store_diagnostics ? 3 : 1));
      double* dst = out + ((size_t)c * nrows + r) * (7 + K);
      std::memcpy(dst, src, 7 * sizeof(double));
      if (constrained) oracle_write_array(&e->m, src + 7, dst + 7);
      else std::memcpy(dst + 7, src + 7, D * sizeof(double));
    }
  return 0;
}
int b200nuts_fetch_diagnostics(b200nuts_engine* e, int64_t chain0, int64_t nchains, int64_t row0, int64_t nrows, double* out) {
  if (!e->cfg.store_diagnostics) { e->err = "no diagnostics are stored"; return B200NUTS_USAGE; }
  const int ncol = 7 + 3 * e->m.D;
  for (int64_t c = 0; c < nchains; c++)
    std::memcpy(out + (size_t)c * nrows * ncol, e->draws[chain0 + c].data() + (size_t)row0 * ncol, sizeof(double) * (size_t)nrows * ncol);
  return 0;
}
int64_t b200nuts_rows_saved(const b200nuts_engine* e) { return e->rows_per_chain; }
int b200nuts_get_adaptation(b200nuts_engine* e, double* stepsize, double* inv_metric) {
  const size_t md = e->cfg.metric == 2 ? (size_t)e->m.D * e->m.D : (size_t)e->m.D;
  for (int k = 0; k < e->cfg.num_chains; k++) {
    if (stepsize) stepsize[k] = e->adapt[k][0];
    if (inv_metric) std::memcpy(inv_metric + (size_t)k * md, e->adapt[k].data() + 1, md * sizeof(double));
  }
  return 0;
}
int b200nuts_get_chain_status(b200nuts_engine* e, int32_t* err, int32_t* iter, int64_t* n_grad) {
  for (int k = 0; k < e->cfg.num_chains; k++) {
    if (err) err[k] = e->rc[k];
    if (iter) iter[k] = e->iters_done;
    if (n_grad) n_grad[k] = e->ngrad[k];
  }
  return 0;
}
int b200nuts_get_timing(b200nuts_engine*, double* s, int64_t* l, double* g) { if (s) *s = 0; if (l) *l = 0; if (g) *g = 0; return 0; }
int b200nuts_summary_num_stats(void) { return 10; }
int b200nuts_summary(b200nuts_engine* e, int64_t, int64_t, int, double*) { e->err = "mock: no summary"; return B200NUTS_SOFTWARE; }
int b200nuts_summarize_draws(const double*, int64_t, int64_t, int64_t, int, double*, char*, int) { return B200NUTS_SOFTWARE; }
int b200nuts_set_allreduce(b200nuts_engine*, b200nuts_allreduce_fn, void*) { return B200NUTS_SOFTWARE; }
int b200nuts_nccl_get_unique_id(void*) { return B200NUTS_SOFTWARE; }
int b200nuts_measure_fp64_fma_peak(int, double*) { return B200NUTS_SOFTWARE; }
int b200nuts_nccl_init(b200nuts_engine*, const void*, int, int) { return B200NUTS_SOFTWARE; }
int b200nuts_measure_allreduce(b200nuts_engine*, int, double*, int64_t*) { return B200NUTS_SOFTWARE; }
int b200nuts_transform_inits(b200nuts_engine*, const double*, int64_t, double*) { return B200NUTS_SOFTWARE; }
int b200nuts_register_model(const b200nuts_model_vtable*) { return -B200NUTS_SOFTWARE; }
int b200nuts_load_model_library(const char*) { g_err = "mock: no plug-ins"; return -B200NUTS_SOFTWARE; }
int b200nuts_find_model(const char* name) {
  static const char* const k[] = {"std_normal", "bernoulli", "eight_schools_nc", "logistic", "irt_2pl", "poisson_glm"};
  for (int i = 0; i < 6; i++) if (!std::strcmp(k[i], name)) return i;
  return -1;
}
int b200nuts_output_name(b200nuts_engine* e, int64_t k, char* buf, int buflen) {  // same names as the engine / host model objects
  if (!e->ready) { int rc = finalize(e); if (rc) return rc; }
  const long long N = e->m.N, K = e->m.K, i = k;
  auto vec = [&](const char* n, long long j) { std::snprintf(buf, (size_t)buflen, "%s.%lld", n, j + 1); };
  auto sc = [&](const char* n) { std::snprintf(buf, (size_t)buflen, "%s", n); };
  switch (e->m.kind) {
    case M_STD_NORMAL: vec("x", i); break;
    case M_BERNOULLI: sc("theta"); break;
    case M_EIGHT_SCHOOLS:
      if (i == 0) sc("mu"); else if (i == 1) sc("tau"); else if (i < 2 + N) vec("theta_tilde", i - 2); else vec("theta", i - 2 - N);
      break;
    case M_LOGISTIC:
    case M_POISSON:
      if (i == 0) sc("alpha"); else vec("beta", i - 1);
      break;
    default:
      if (i == 0) sc("sigma_theta"); else if (i < 1 + N) vec("theta", i - 1); else if (i == 1 + N) sc("sigma_a");
      else if (i < 2 + N + K) vec("a", i - 2 - N); else if (i == 2 + N + K) sc("mu_b"); else if (i == 3 + N + K) sc("sigma_b");
      else vec("b", i - 4 - N - K);
  }
  return 0;
}

}  // extern "C"
