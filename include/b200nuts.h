/*
 * b200nuts.h -- C ABI of libb200nuts.so, the B200 (sm_100a) NUTS engine.
 *
 * Drop-in boundary: the reference has no FFI on this path; its one seam is the service call
 *
 *     stan::services::sample::hmc_nuts_diag_e_adapt(model, num_chains, init, init_inv_metric, random_seed,
 *         init_chain_id, init_radius, num_warmup, num_samples, num_thin, save_warmup, refresh, stepsize,
 *         stepsize_jitter, max_depth, delta, gamma, kappa, t0, init_buffer, term_buffer, window, interrupt,
 *         logger, init_writer, sample_writer, diagnostic_writer, metric_writer)
 *                                   (stan/src/stan/services/sample/hmc_nuts_diag_e_adapt.hpp:328-343, multi-chain :331-407)
 *
 * made from cmdstan::command() (src/cmdstan/command.hpp:673-688).  The host shim
 * b200::services::sample::hmc_nuts_diag_e_adapt (cmdstan_b200/host/service.hpp) keeps that parameter list and
 * forwards to the functions below; INTEGRATION.md shows the one-line change in command.hpp.
 *
 * Conventions: every function returns a stan::services::error_codes value
 * (stan/src/stan/services/error_codes.hpp: OK 0, USAGE 64, DATAERR 65, NOINPUT 66, SOFTWARE 70, CONFIG 78); the
 * message is available from b200nuts_last_error().  All pointers are HOST pointers; inputs are copied before the call
 * returns, outputs are caller-allocated.  An engine handle owns all device memory and is driven by one host thread.
 * There is no CPU fallback: without a CUDA device every call fails with SOFTWARE.
 */
#ifndef B200NUTS_H
#define B200NUTS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200NUTS_OK 0
#define B200NUTS_USAGE 64
#define B200NUTS_DATAERR 65
#define B200NUTS_NOINPUT 66
#define B200NUTS_SOFTWARE 70
#define B200NUTS_CONFIG 78

/* model family implemented on the device (closed-form log density + gradient) */
enum b200nuts_model {
  B200NUTS_STD_NORMAL = 0,       /* x[D] ~ normal(0,1)                       data: int "D" */
  B200NUTS_BERNOULLI = 1,        /* examples/bernoulli/bernoulli.stan        data: int "y"[N] */
  B200NUTS_EIGHT_SCHOOLS_NC = 2, /* non-centred eight schools                data: f64 "y"[J], f64 "sigma"[J] */
  B200NUTS_LOGISTIC_GLM = 3,     /* bernoulli_logit_glm, N(0,2.5)/N(0,1)     data: f64 "X"[N][K] row-major, int "y"[N] */
  B200NUTS_IRT_2PL = 4,          /* hierarchical IRT 2PL                     data: i8  "y"[I][J] row-major */
  B200NUTS_POISSON_GLM = 5       /* poisson_log_glm, N(0,1) priors           data: f64 "X"[N][K] row-major, int "y"[N] */
};

/* which engine runs the chains */
enum b200nuts_engine_kind {
  B200NUTS_ENGINE_AUTO = 0,     /* fused for models 0-2, lock-step otherwise */
  B200NUTS_ENGINE_FUSED = 1,    /* one persistent warp per chain, model gradient inlined (models 0-2 only) */
  B200NUTS_ENGINE_LOCKSTEP = 2  /* resumable chains + one cross-chain gradient kernel per leapfrog */
};

/* Arguments of hmc_nuts_diag_e_adapt (same names, same meaning, same defaults as src/cmdstan/arguments/arg_*.hpp). */
typedef struct b200nuts_config {
  int32_t model;            /* enum b200nuts_model */
  int32_t num_chains;       /* chains run by THIS engine (one engine per GPU / rank) */
  uint32_t random_seed;     /* random seed=           (create_rng(seed, chain)) */
  uint32_t init_chain_id;   /* id= ; chain c uses stream id init_chain_id + c */
  double init_radius;       /* init=  (default 2; 0 => all-zero inits) */
  int32_t num_warmup;       /* default 1000 */
  int32_t num_samples;      /* default 1000 */
  int32_t num_thin;         /* default 1 */
  int32_t save_warmup;      /* default 0 */
  double stepsize;          /* default 1 */
  double stepsize_jitter;   /* default 0 */
  int32_t max_depth;        /* default 10 (1..16) */
  double delta;             /* default 0.8 */
  double gamma;             /* default 0.05 */
  double kappa;             /* default 0.75 */
  double t0;                /* default 10 */
  uint32_t init_buffer;     /* default 75 */
  uint32_t term_buffer;     /* default 50 */
  uint32_t window;          /* default 25 */
  int32_t adapt_engaged;    /* default 1; 0 => services::sample::hmc_nuts_diag_e (no adaptation, no init_stepsize) */
  int32_t engine;           /* enum b200nuts_engine_kind */
  int32_t device;           /* CUDA device ordinal */
  int32_t store_draws;      /* 1: keep every saved draw in HBM for b200nuts_fetch_draws (default); 0: keep none */
  /* row sharding (mode 2): this engine holds rows [row_begin, row_begin + N) of a GLM whose
     partial (lp, gradient) sums are combined across `row_shards` engines by the all-reduce hook below. */
  int32_t row_shards;       /* default 1 */
  int32_t row_shard_rank;   /* default 0 */
  int32_t metric;           /* 0 diag_e (default): hmc_nuts_diag_e[_adapt]; 1 unit_e: hmc_nuts_unit_e[_adapt]
                               (stan/src/stan/services/sample/hmc_nuts_unit_e_adapt.hpp; the metric stays the identity);
                               2 dense_e: hmc_nuts_dense_e[_adapt] (stan/src/stan/services/sample/hmc_nuts_dense_e_adapt.hpp,
                               stan/mcmc/hmc/hamiltonians/dense_e_metric.hpp, stan/mcmc/covar_adaptation.hpp): a D x D
                               inverse metric per chain, covariance learned in the adaptation windows */
  int32_t store_diagnostics; /* default 0; 1: every saved draw also keeps the sample's momentum and the gradient of the potential
                               (what the reference sends to `output diagnostic_file=`: mcmc_writer::write_diagnostic_params,
                               stan/src/stan/services/util/mcmc_writer.hpp:156-176), fetched with b200nuts_fetch_diagnostics */
  int32_t static_hmc;       /* default 0; 1: static HMC instead of NUTS -- services hmc_static_{diag,unit,dense}_e[_adapt]
                               (stan/src/stan/services/sample/hmc_static_*.hpp, stan/mcmc/hmc/static/base_static_hmc.hpp): every
                               transition runs L = max(1, (int)(int_time / stepsize)) leapfrogs and a Metropolis test.  Saved rows
                               keep seven leading columns: lp__, accept_stat__, stepsize__, L, 1, 0, energy__ (CmdStan's writer
                               prints stepsize__, int_time__, energy__) */
  double int_time;          /* default 2 pi (arguments/arg_static.hpp) */
} b200nuts_config;

typedef struct b200nuts_engine b200nuts_engine;

/* fills *cfg with the reference's defaults */
void b200nuts_default_config(b200nuts_config* cfg);

/* create / destroy.  The model data must be set before b200nuts_init_chains. */
int b200nuts_create(const b200nuts_config* cfg, b200nuts_engine** out);
void b200nuts_destroy(b200nuts_engine* e);
const char* b200nuts_last_error(const b200nuts_engine* e); /* e may be NULL: error of the last failed create */

/* model data, by name (see enum b200nuts_model); dims row-major */
int b200nuts_set_data_f64(b200nuts_engine* e, const char* name, const double* host, const int64_t* dims, int nd);
int b200nuts_set_data_i32(b200nuts_engine* e, const char* name, const int32_t* host, const int64_t* dims, int nd);
int b200nuts_set_data_i8(b200nuts_engine* e, const char* name, const int8_t* host, const int64_t* dims, int nd);

/* dimension of the unconstrained space / number of write_array output columns (valid after the data is set) */
int b200nuts_num_params(const b200nuts_engine* e, int64_t* num_unconstrained, int64_t* num_outputs);

/* optional user inits q0[num_chains][D] (unconstrained scale) and inverse metric [num_chains][D] (dense_e:
   [num_chains][D][D], symmetric positive definite, else CONFIG "Inverse Euclidean metric not positive definite.");
   NULL => reference behaviour (uniform(-R,R) draws from each chain's own stream; unit metric).  Partial inits: an element of
   q0 that is a quiet NaN is "not given" and takes the value the reference's random context would give it (the low two payload
   bits name the parameter's transform for the constrain/unconstrain round trip: 0 none, 1 lower bound 0, 2 unit interval) */
int b200nuts_set_inits(b200nuts_engine* e, const double* q0);
int b200nuts_set_inv_metric(b200nuts_engine* e, const double* inv_metric);

/* seeds the per-chain MIXMAX streams and allocates all chain state; must precede b200nuts_run */
int b200nuts_init_chains(b200nuts_engine* e);

/* stan::model::log_prob_grad<propto=true, jacobian>(model, q) for n points (gate 1 of the parity checks) */
int b200nuts_log_prob_grad(b200nuts_engine* e, const double* q, int64_t n, int jacobian, double* lp, double* grad);

/* advance every chain by n_iter transitions (warm-up transitions first, then sampling ones; initialisation and
   init_stepsize happen inside the first call).  Blocks until all chains are idle. */
int b200nuts_run(b200nuts_engine* e, int n_iter);

/* saved draws of chains [chain0, chain0+nchains), rows [row0, row0+nrows): out[nchains][nrows][7 + K] with columns
   lp__, accept_stat__, stepsize__, treedepth__, n_leapfrog__, divergent__, energy__, then model.write_array
   (constrained = 1, K = num_outputs) or the raw unconstrained q (constrained = 0, K = D) */
int b200nuts_fetch_draws(b200nuts_engine* e, int64_t chain0, int64_t nchains, int64_t row0, int64_t nrows, int constrained,
                         double* out);
/* rows of the diagnostic output (needs store_diagnostics = 1): out[nchains][nrows][7 + 3 D] = the seven sampler columns, the
   unconstrained parameters q, the momentum p and the gradient of the potential g = -grad lp of each saved draw */
int b200nuts_fetch_diagnostics(b200nuts_engine* e, int64_t chain0, int64_t nchains, int64_t row0, int64_t nrows, double* out);
int64_t b200nuts_rows_saved(const b200nuts_engine* e); /* rows every chain has saved so far (min over chains) */

/* On-device posterior summary of the saved rows [row0, row0+nrows) of all chains of this engine: the table the reference's
   `stansummary` prints (src/cmdstan/stansummary_helper.hpp:405-428), computed from the draws where they live in HBM
   (stan/src/stan/analyze/mcmc/{split_chains,rank_normalization,autocovariance,ess,rhat,mcse}.hpp, stan/mcmc/chainset.hpp).
   out[7 + K][b200nuts_summary_num_stats()] row-major, one row per output column (the 7 sampler columns first), columns
   Mean, MCSE, StdDev, MAD, 5%, 50%, 95%, ESS_bulk, ESS_tail, R_hat. */
int b200nuts_summary_num_stats(void);
int b200nuts_summary(b200nuts_engine* e, int64_t row0, int64_t nrows, int constrained, double* out);
/* the same table for draws held by the caller (e.g. read back from Stan CSV files, as bin/stansummary does):
   draws[chains][rows][ncols] row-major host memory -> out[ncols][num_stats]; the message of a failure is copied to errbuf */
int b200nuts_summarize_draws(const double* draws, int64_t chains, int64_t rows, int64_t ncols, int device, double* out,
                             char* errbuf, int errlen);

/* adapted step size [num_chains] and inverse metric [num_chains][D] ("# Step size", "# Diagonal elements ...");
   dense_e: [num_chains][D][D] ("# Elements of inverse mass matrix:") */
int b200nuts_get_adaptation(b200nuts_engine* e, double* stepsize, double* inv_metric);
/* per-chain status: error code (0 ok, 78 init failed, 70 step-size / metric failure), iterations done, leapfrog
   gradients consumed; any pointer may be NULL */
int b200nuts_get_chain_status(b200nuts_engine* e, int32_t* err, int32_t* iter, int64_t* n_grad);
/* device time of the last b200nuts_run in seconds, number of kernels it launched, and the time of its dominant
   kernel family (gradient stage) measured with CUDA events on the launching stream */
int b200nuts_get_timing(b200nuts_engine* e, double* run_seconds, int64_t* kernel_launches, double* grad_kernel_seconds);

/* Measured FP64 FMA rate of `device` in TFLOP/s (a register-resident chain of independent DFMAs on every SM): the roofline
   denominator of the kernels bound by the scalar FP64 pipe (IRT), SURVEY.md section 8d "must be measured on the box". */
int b200nuts_measure_fp64_fma_peak(int device, double* tflops);

/* Row-sharded GLMs: called on the engine's stream position after the partial gradient kernel of every leapfrog with a
   DEVICE buffer of `count` doubles that must be summed element-wise across all shards, identically on every shard
   (e.g. ncclAllReduce on `cuda_stream`).  Must be set when row_shards > 1. */
typedef int (*b200nuts_allreduce_fn)(void* device_buf, int64_t count, void* cuda_stream, void* user);
int b200nuts_set_allreduce(b200nuts_engine* e, b200nuts_allreduce_fn fn, void* user);

/* Row sharding over NCCL (mode 2: "one ncclAllReduce(sum, double, (1+D) C) per leapfrog").  NCCL is
   loaded at run time (dlopen of libnccl.so.2, or the path in the environment variable B200NUTS_NCCL_LIB).  Rank 0 obtains a
   128-byte unique id and hands it to every rank by any channel; each rank then creates its communicator on the engine's
   device; this installs the all-reduce hook above with ncclAllReduce on the engine's stream. */
int b200nuts_nccl_get_unique_id(void* id128);
int b200nuts_nccl_init(b200nuts_engine* e, const void* id128, int nranks, int rank);
/* Cost of the exchange step of a row-sharded engine: enqueues `iters` all-reduces of exactly the per-leapfrog payload
   ((1 + D padded to 8) x chains doubles) back to back on the engine's stream through the installed hook and returns their
   average duration in microseconds, timed with CUDA events (collective: every shard must call it with the same `iters`).
   The payload buffer is zeroed first and is scratch between gradient stages, so chains are unaffected. */
int b200nuts_measure_allreduce(b200nuts_engine* e, int iters, double* usec_per_allreduce, int64_t* payload_bytes);

#ifdef __cplusplus
}
#endif
#endif
