/* TEST INFRASTRUCTURE ONLY (see nuts_oracle.c): the oracle's public types and entry points, shared by nuts_oracle.c and the
 * test-only mock of the C ABI (tests/host_emul/mock_b200nuts.cpp). */
#ifndef NUTS_ORACLE_H
#define NUTS_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { M_STD_NORMAL = 0, M_BERNOULLI = 1, M_EIGHT_SCHOOLS = 2, M_LOGISTIC = 3, M_IRT2PL = 4, M_POISSON = 5,
       M_MOCK_FLAT = 6 /* log_prob == 0: the mock_model of ST/unit/mcmc/hmc/mock_hmc.hpp:13-35, KATs only */ };

typedef struct {
  int kind;
  int D;              /* unconstrained dimension */
  int N;              /* rows / observations (bernoulli N, eight schools J, GLM N, IRT persons I) */
  int K;              /* covariates (GLM) / items J (IRT) */
  const double* X;    /* GLM: N x K row-major */
  const int* yi;      /* bernoulli / GLM outcomes */
  const double* yr;   /* eight schools y */
  const double* sigma;/* eight schools sigma */
  const signed char* y8; /* IRT: I x J row-major 0/1 */
} model_t;

typedef struct {
  uint32_t seed, chain_id;
  int num_warmup, num_samples, num_thin, save_warmup;
  double stepsize, stepsize_jitter;
  int max_depth;
  double delta, gamma, kappa, t0;
  unsigned init_buffer, term_buffer, window;
  double init_radius;
  int adapt_engaged;
  int unit_e; /* 1: adapt_unit_e_nuts / unit_e_nuts (S/mcmc/hmc/nuts/adapt_unit_e_nuts.hpp:24-35): metric frozen at the identity,
                 step-size adaptation only -- arithmetically diag_e with M^-1 = 1 and no variance windows */
  int dense_e; /* 1: adapt_dense_e_nuts / dense_e_nuts (S/mcmc/hmc/nuts/adapt_dense_e_nuts.hpp:20-50): dense inverse metric
                  (S/mcmc/hmc/hamiltonians/dense_e_metric.hpp:21-58, dense_e_point.hpp), covariance windows
                  (S/mcmc/covar_adaptation.hpp:17-45, M/prim/fun/welford_covar_estimator.hpp:10-50) */
  int diag;    /* 1: rows carry the momentum and the potential's gradient of the sample after the parameters (7 + 3 D columns):
                  what mcmc_writer::write_diagnostic_params sends to the diagnostic writer (services/util/mcmc_writer.hpp:156-176,
                  base_hmc::get_sampler_diagnostics -> ps_point::get_params: q, p, g) */
  int static_hmc;  /* 1: static HMC (S/mcmc/hmc/static/base_static_hmc.hpp:41-70, adapt_diag_e_static_hmc.hpp:27-50 and its unit_e /
                      dense_e siblings): L = max(1, (int)(int_time / nominal stepsize)) leapfrogs, Metropolis accept.  Rows keep seven
                      leading columns: lp__, accept_stat__, stepsize__, L, 1, 0, energy__ (the writer prints stepsize__, int_time__,
                      energy__) */
  double int_time;
} oracle_config_t;

double oracle_lp_grad(const model_t* m, const double* q, double* g);
int oracle_num_outputs(const model_t* m);
void oracle_write_array(const model_t* m, const double* q, double* out);
int oracle_run_chain(const model_t* model, const oracle_config_t* c, const double* q0, const double* inv_metric0, double* draws,
                     double* adapt_out, double* q_init_out, long* grad_evals_out);

/* the reference's sampler unit tests restated: base_nuts_test.cpp:139-195, 335-372; unit_e_nuts_test.cpp:10-95 */
void oracle_nuts_transition_kat(const model_t* model, uint32_t seed, uint32_t chain, const double* q0, const double* p0, double eps,
                                int max_depth, double* out);
void oracle_nuts_build_tree_kat(const model_t* model, uint32_t seed, uint32_t chain, const double* q0, const double* p0, double eps, int depth,
                                double sign, double H0, int h0_relative, double* out);

int oracle_learn_covariance_kat(int D, const double* q, int n_learn, unsigned num_warmup, unsigned init_buffer, unsigned term_buffer,
                                unsigned window, double* covar_out);
int oracle_criterion_kat(const double* p_sharp_minus, const double* p_sharp_plus, const double* rho, int n);

#ifdef __cplusplus
}
#endif
#endif
