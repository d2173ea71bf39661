/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle, never linked into / imported by the product.
 *
 * Plain-C scalar restatement of the reference's NUTS diag_e adaptive sampling path
 * (CmdStan 2.36 / Stan 2.36), following the reference's own *recursive* formulation and RNG
 * consumption order.  The CUDA engine uses a different (iterative, warp-parallel, resumable) formulation;
 * agreement between the two and with oracle/_ref (the compiled unmodified reference) is the parity gate.
 *
 * Pinned against: Appendix-B RNG known-answer vectors, ST/unit/mcmc KATs, and CSV output of
 * oracle/_ref at sig_figs=18 (tests/test_oracle_*.py; fixtures under tests/golden/).
 *
 * Citations (abbreviations as SURVEY.md: S/ = stan/src/stan, M/ = stan_math/stan/math, B/ = boost):
 *   RNG            B/random/mixmax.hpp:104-112,194-208,210-301; S/services/util/create_rng.hpp:28-34
 *   uniform_01     B/random/uniform_01.hpp:137-156
 *   uniform_real   B/random/uniform_real_distribution.hpp:47-64
 *   normal         B/random/normal_distribution.hpp:120-201, detail/int_float_pair.hpp:34-98,
 *                  exponential_distribution.hpp:181-218
 *   initialize     S/services/util/initialize.hpp:70-260, S/io/random_var_context.hpp:43-66
 *   init_stepsize  S/mcmc/hmc/base_hmc.hpp:93-161
 *   transition     S/mcmc/hmc/nuts/base_nuts.hpp:78-204 ; build_tree :247-352 ; criterion :222-226
 *   leapfrog       S/mcmc/hmc/integrators/expl_leapfrog.hpp:16-32, base_leapfrog.hpp:17-22
 *   metric         S/mcmc/hmc/hamiltonians/diag_e_metric.hpp:20-50
 *   adaptation     S/mcmc/stepsize_adaptation.hpp:49-73, var_adaptation.hpp:17-46,
 *                  windowed_adaptation.hpp:23-110, M/prim/fun/welford_var_estimator.hpp:10-45,
 *                  S/mcmc/hmc/nuts/adapt_diag_e_nuts.hpp:25-48
 *   log_sum_exp    M/prim/fun/log_sum_exp.hpp:51-64, log1p_exp.hpp:45-52
 *   densities      see each lp_grad_* below
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ RNG ---------------------------- */
#define MM_N 17
#define M61 2305843009213693951ULL

typedef struct {
  uint64_t V[MM_N];
  uint64_t sumtot;
  int counter;
} mixmax_t;

static const uint64_t kSkip[128][MM_N] = {
#include "mixmax_skip_n17.inc"
};
#define B200_TABLE static const
#include "ziggurat_tables.inc"

static inline uint64_t mod61(uint64_t k) { return (k & M61) + (k >> 61); }
static inline uint64_t mulwu(uint64_t k) { return ((k << 36) & M61) ^ (k >> 25); }

static uint64_t mm_iterate(uint64_t* Y, uint64_t sum_old) {
  uint64_t tempP = 0, tempV = sum_old;
  Y[0] = tempV;
  uint64_t sumtot = Y[0], ovflow = 0;
  for (int i = 1; i < MM_N; i++) {
    uint64_t tempPO = mulwu(tempP);
    tempV = tempV + tempPO;
    tempP = mod61(tempP + Y[i]);
    tempV = mod61(tempV + tempP);
    Y[i] = tempV;
    sumtot += tempV;
    if (sumtot < tempV) ovflow++;
  }
  return mod61(mod61(sumtot) + (ovflow << 3));
}

static uint64_t fmodmul61(uint64_t cum, uint64_t s, uint64_t a) {
  const uint64_t MASK32 = 0xFFFFFFFFULL;
  uint64_t o = s * a, ph = s >> 32, pl = s & MASK32, ah = a >> 32, al = a & MASK32;
  o = (o & M61) + ((ph * ah) << 3) + ((ah * pl + al * ph + ((al * pl) >> 32)) >> 29);
  o += cum;
  o = (o & M61) + (o >> 61);
  return o;
}

/* stan::services::util::create_rng(seed, chain) == mixmax(clusterID=0, machineID=1, runID=seed, streamID=chain) */
void oracle_rng_seed(mixmax_t* s, uint32_t seed, uint32_t chain) {
  uint32_t id[4] = {chain, seed, 1u, 0u};
  uint64_t Y[MM_N], cum[MM_N], sumtot = 0;
  memset(Y, 0, sizeof Y);
  Y[0] = 1;
  for (int i = 0; i < MM_N; i++) sumtot = mod61(sumtot + Y[i]);
  for (int k = 0; k < 4; k++) {
    uint32_t v = id[k];
    int r = 0;
    while (v) {
      if (v & 1) {
        const uint64_t* row = kSkip[r + 32 * k];
        for (int i = 0; i < MM_N; i++) cum[i] = 0;
        for (int j = 0; j < MM_N; j++) {
          uint64_t coeff = row[j];
          for (int i = 0; i < MM_N; i++) cum[i] = fmodmul61(cum[i], coeff, Y[i]);
          sumtot = mm_iterate(Y, sumtot);
        }
        sumtot = 0;
        for (int i = 0; i < MM_N; i++) {
          Y[i] = cum[i];
          sumtot = mod61(sumtot + cum[i]);
        }
      }
      v >>= 1;
      r++;
    }
  }
  sumtot = 0;
  for (int i = 0; i < MM_N; i++) {
    s->V[i] = Y[i];
    sumtot = mod61(sumtot + Y[i]);
  }
  s->sumtot = sumtot;
  s->counter = 1;
}

uint64_t oracle_rng_next(mixmax_t* s) {
  if (s->counter <= MM_N - 1) return s->V[s->counter++];
  s->sumtot = mm_iterate(s->V, s->sumtot);
  s->counter = 2;
  return s->V[1];
}

#define TWO_M61 4.336808689942017736029811203479766845703125e-19 /* 2^-61 */
#define TWO_M53 1.1102230246251565404236316680908203125e-16      /* 2^-53 */

double oracle_uniform01(mixmax_t* s) {
  for (;;) {
    double r = (double)oracle_rng_next(s) * TWO_M61;
    if (r < 1.0) return r;
  }
}

double oracle_uniform_real(mixmax_t* s, double a, double b) {
  for (;;) {
    double num = (double)oracle_rng_next(s);
    double div = 2305843009213693952.0; /* double(2^61-1) + 1 */
    double r = num / div * (b - a) + a;
    if (r < b) return r;
  }
}

static double unit_exponential(mixmax_t* s) {
  double shift = 0;
  for (;;) {
    uint64_t u = oracle_rng_next(s);
    int i = (int)(u & 0xFF);
    double x = (double)(u >> 8) * TWO_M53 * kZigExpX[i];
    if (x < kZigExpX[i + 1]) return shift + x;
    if (i == 0) {
      shift += kZigExpX[1];
    } else {
      double y01 = oracle_uniform01(s);
      double y = kZigExpY[i] + y01 * (kZigExpY[i + 1] - kZigExpY[i]);
      double y_above_ubound = (kZigExpX[i] - kZigExpX[i + 1]) * y01 - (kZigExpX[i] - x);
      double y_above_lbound = y - (kZigExpY[i + 1] + (kZigExpX[i + 1] - x) * kZigExpY[i + 1]);
      if (y_above_ubound < 0 && (y_above_lbound < 0 || y < exp(-x))) return x + shift;
    }
  }
}

double oracle_normal(mixmax_t* s) {
  for (;;) {
    uint64_t u = oracle_rng_next(s);
    int bucket = (int)(u & 0xFF);
    double r = (double)(u >> 8) * TWO_M53;
    int sign = (bucket & 1) * 2 - 1;
    int i = bucket >> 1;
    double x = r * kZigNormX[i];
    if (x < kZigNormX[i + 1]) return x * sign;
    if (i == 0) {
      const double tail_start = kZigNormX[1];
      for (;;) {
        double tx = unit_exponential(s) / tail_start;
        double ty = unit_exponential(s);
        if (2 * ty > tx * tx) return (tx + tail_start) * sign;
      }
    }
    double y01 = oracle_uniform01(s);
    double y = kZigNormY[i] + y01 * (kZigNormY[i + 1] - kZigNormY[i]);
    double y_above_ubound, y_above_lbound;
    if (kZigNormX[i] >= 1) {
      y_above_ubound = (kZigNormX[i] - kZigNormX[i + 1]) * y01 - (kZigNormX[i] - x);
      y_above_lbound = y - (kZigNormY[i] + (kZigNormX[i] - x) * kZigNormY[i] * kZigNormX[i]);
    } else {
      y_above_lbound = (kZigNormX[i] - kZigNormX[i + 1]) * y01 - (kZigNormX[i] - x);
      y_above_ubound = y - (kZigNormY[i] + (kZigNormX[i] - x) * kZigNormY[i] * kZigNormX[i]);
    }
    if (y_above_ubound < 0 && (y_above_lbound < 0 || y < exp(-(x * x / 2)))) return x * sign;
  }
}

/* KAT helpers for the python tests */
void oracle_rng_kat(uint32_t seed, uint32_t chain, int kind, int n, double* out_f, uint64_t* out_u) {
  mixmax_t s;
  oracle_rng_seed(&s, seed, chain);
  for (int i = 0; i < n; i++) {
    if (kind == 0) out_u[i] = oracle_rng_next(&s);
    else if (kind == 1) out_f[i] = oracle_uniform01(&s);
    else if (kind == 2) out_f[i] = oracle_normal(&s);
    else out_f[i] = oracle_uniform_real(&s, -2.0, 2.0);
  }
}

/* ------------------------------------------------------------------ models ------------------------- */
#include "nuts_oracle.h"


static inline double log1p_exp(double a) { return a > 0.0 ? a + log1p(exp(-a)) : log1p(exp(a)); }

static inline double inv_logit(double u) { /* M/prim/fun/inv_logit.hpp:51-60 */
  if (u < 0) {
    double e = exp(u);
    if (u < -36.04365338911715) return e; /* LOG_EPSILON */
    return e / (1 + e);
  }
  return 1 / (1 + exp(-u));
}

/* x ~ normal(0,1), propto: -0.5 sum x^2 (M/prim/prob/normal_lpdf.hpp:83-89) */
static double lp_grad_std_normal(const model_t* m, const double* q, double* g) {
  double s = 0;
  for (int i = 0; i < m->D; i++) {
    s += q[i] * q[i];
    g[i] = -q[i];
  }
  return -0.5 * s;
}

/* c1: theta = inv_logit(u); Jacobian M/rev/constraint/lub_constrain.hpp:100-148 (prim form: -|u| - 2 log1p_exp(-|u|));
 * bernoulli_lpmf.hpp:34-114 (sufficient statistic sum(y)); beta_lpdf(1,1) contributes 0*log(theta) + 0*log1m(theta). */
static double lp_grad_bernoulli(const model_t* m, const double* q, double* g) {
  double u = q[0];
  int S = 0;
  for (int i = 0; i < m->N; i++) S += m->yi[i];
  int N = m->N;
  double th = inv_logit(u);
  double au = fabs(u);
  double lp = -au - 2.0 * log1p_exp(-au);
  double lth = log(th), l1m = log1p(-th);
  lp += 0.0 * lth + 0.0 * l1m;
  double dth;
  if (S == N) {
    lp += N * lth;
    dth = N / th;
  } else if (S == 0) {
    lp += N * l1m;
    dth = N / (th - 1);
  } else {
    lp += S * lth + (N - S) * l1m;
    dth = S / th + (N - S) / (th - 1);
  }
  g[0] = dth * th * (1 - th) + (1 - 2 * th);
  return lp;
}

/* c2: SURVEY.md Appendix E; normal_lpdf.hpp:44-105, cauchy_lpdf.hpp:80-86, lb_constrain (lp += u) */
static double lp_grad_eight_schools(const model_t* m, const double* q, double* g) {
  int J = m->N;
  double mu = q[0], ell = q[1], tau = exp(ell);
  double lp = ell;
  double z = mu / 5.0;
  lp += -0.5 * z * z;
  double t5 = tau / 5.0;
  lp += -log1p(t5 * t5);
  double ss = 0, sr = 0, srt = 0, sq = 0;
  for (int j = 0; j < J; j++) {
    double tt = q[2 + j];
    double theta = mu + tau * tt;
    double zs = (m->yr[j] - theta) / m->sigma[j];
    double r = zs / m->sigma[j];
    ss += tt * tt;
    sq += zs * zs;
    sr += r;
    srt += r * tt;
    g[2 + j] = -tt + tau * r;
  }
  lp += -0.5 * ss - 0.5 * sq;
  g[0] = -mu / 25.0 + sr;
  g[1] = 1.0 - 2.0 * t5 * t5 / (1.0 + t5 * t5) + tau * srt;
  return lp;
}

/* c3: bernoulli_logit_glm_lpmf.hpp:51-167 (cutoff +-20), priors alpha ~ N(0,2.5), beta ~ N(0,1) */
static double lp_grad_logistic(const model_t* m, const double* q, double* g) {
  int N = m->N, K = m->K;
  double alpha = q[0];
  const double* beta = q + 1;
  double lp = 0, sd = 0;
  for (int j = 0; j < K; j++) g[1 + j] = 0;
  for (int i = 0; i < N; i++) {
    const double* x = m->X + (size_t)i * K;
    double eta = 0;
    for (int j = 0; j < K; j++) eta += x[j] * beta[j];
    double sgn = 2.0 * m->yi[i] - 1.0;
    double t = sgn * (eta + alpha);
    double e = exp(-t), d;
    if (t > 20.0) {
      lp += -e;
      d = -e;
    } else if (t < -20.0) {
      lp += t;
      d = sgn;
    } else {
      lp += -log1p(e);
      d = sgn * e / (e + 1);
    }
    sd += d;
    for (int j = 0; j < K; j++) g[1 + j] += x[j] * d;
  }
  double za = alpha / 2.5, sb = 0;
  for (int j = 0; j < K; j++) {
    sb += beta[j] * beta[j];
    g[1 + j] += -beta[j];
  }
  lp += -0.5 * za * za - 0.5 * sb;
  g[0] = sd - za / 2.5;
  return lp;
}

/* c5: poisson_log_glm_lpmf.hpp:53-163; priors alpha ~ N(0,1), beta ~ N(0,1).  Non-finite sum(d) => domain_error => -inf */
static double lp_grad_poisson(const model_t* m, const double* q, double* g) {
  int N = m->N, K = m->K;
  double alpha = q[0];
  const double* beta = q + 1;
  double lp = 0, sd = 0;
  for (int j = 0; j < K; j++) g[1 + j] = 0;
  for (int i = 0; i < N; i++) {
    const double* x = m->X + (size_t)i * K;
    double eta = 0;
    for (int j = 0; j < K; j++) eta += x[j] * beta[j];
    eta += alpha;
    double ee = exp(eta);
    double d = m->yi[i] - ee;
    lp += m->yi[i] * eta - ee;
    sd += d;
    for (int j = 0; j < K; j++) g[1 + j] += x[j] * d;
  }
  double sb = 0;
  for (int j = 0; j < K; j++) {
    sb += beta[j] * beta[j];
    g[1 + j] += -beta[j];
  }
  lp += -0.5 * alpha * alpha - 0.5 * sb;
  g[0] = sd - alpha;
  if (!isfinite(sd)) return -INFINITY;
  return lp;
}

/* c4: IRT 2PL; order (sigma_theta, theta[I], sigma_a, a[J], mu_b, sigma_b, b[J]); positive ones enter as logs.
 * cauchy_lpdf.hpp:80-86, normal_lpdf.hpp:83-89 (keeps -N log sigma), lognormal_lpdf.hpp:72-80 (keeps N*NEG_LOG_SQRT_TWO_PI),
 * bernoulli_logit_lpmf.hpp:73-89 */
static double lp_grad_irt(const model_t* m, const double* q, double* g) {
  int I = m->N, J = m->K;
  const double NEG_LOG_SQRT_TWO_PI = -0.91893853320467274178;
  double u_st = q[0];
  const double* theta = q + 1;
  double u_sa = q[1 + I];
  const double* u_a = q + 2 + I;
  double mu_b = q[2 + I + J];
  double u_sb = q[3 + I + J];
  const double* b = q + 4 + I + J;
  double* g_theta = g + 1;
  double* g_ua = g + 2 + I;
  double* g_b = g + 4 + I + J;
  double st = exp(u_st), sa = exp(u_sa), sb = exp(u_sb);
  double lp = u_st + u_sa + u_sb; /* Jacobians of the three scalar lower bounds */
  if (!(st > 0 && isfinite(st) && sa > 0 && isfinite(sa) && sb > 0 && isfinite(sb))) {
    for (int i = 0; i < m->D; i++) g[i] = 0;
    return -INFINITY; /* check_positive_finite => domain_error */
  }
  /* sigma_theta ~ cauchy(0,2) */
  double h = st / 2.0;
  lp += -log1p(h * h);
  double g_st = -2.0 * h * h / (1.0 + h * h); /* d/du of -log1p((e^u/2)^2) */
  /* theta ~ normal(0, sigma_theta) */
  double s2 = 0;
  for (int i = 0; i < I; i++) {
    double z = theta[i] / st;
    s2 += z * z;
    g_theta[i] = -z / st;
  }
  lp += -0.5 * s2 - I * u_st;
  g_st += s2 - I;
  /* sigma_a ~ cauchy(0,2) */
  h = sa / 2.0;
  lp += -log1p(h * h);
  double g_sa = -2.0 * h * h / (1.0 + h * h);
  /* a ~ lognormal(0, sigma_a), a = exp(u_a), Jacobian sum(u_a) */
  double s2a = 0, sua = 0;
  for (int j = 0; j < J; j++) {
    double z = u_a[j] / sa;
    s2a += z * z;
    sua += u_a[j];
    g_ua[j] = -z / sa - 1.0 + 1.0; /* d/du_a[-0.5 (u/sa)^2 - u] + Jacobian (+1) */
  }
  lp += J * NEG_LOG_SQRT_TWO_PI - 0.5 * s2a - J * u_sa - sua + sua;
  g_sa += s2a - J;
  /* mu_b ~ normal(0,5) */
  double zb = mu_b / 5.0;
  lp += -0.5 * zb * zb;
  double g_mub = -zb / 5.0;
  /* sigma_b ~ cauchy(0,2) */
  h = sb / 2.0;
  lp += -log1p(h * h);
  double g_sb = -2.0 * h * h / (1.0 + h * h);
  /* b ~ normal(mu_b, sigma_b) */
  double s2b = 0;
  for (int j = 0; j < J; j++) {
    double z = (b[j] - mu_b) / sb;
    s2b += z * z;
    g_b[j] = -z / sb;
    g_mub += z / sb;
  }
  lp += -0.5 * s2b - J * u_sb;
  g_sb += s2b - J;
  /* likelihood */
  for (int j = 0; j < J; j++) {
    double a = exp(u_a[j]);
    double ga = 0, gb = 0;
    for (int i = 0; i < I; i++) {
      double sgn = 2.0 * m->y8[(size_t)i * J + j] - 1.0;
      double diff = theta[i] - b[j];
      double t = sgn * (a * diff);
      double e = exp(-t), d;
      if (t > 20.0) {
        lp += -e;
        d = -e;
      } else if (t < -20.0) {
        lp += t;
        d = sgn;
      } else {
        lp += -log1p(e);
        d = sgn * e / (e + 1);
      }
      g_theta[i] += d * a;
      gb -= d * a;
      ga += d * diff;
    }
    g_b[j] += gb;
    g_ua[j] += ga * a;
  }
  g[0] = g_st + 1.0; /* + Jacobian d(u)/du */
  g[1 + I] = g_sa + 1.0;
  g[2 + I + J] = g_mub;
  g[3 + I + J] = g_sb + 1.0;
  return lp;
}

double oracle_lp_grad(const model_t* m, const double* q, double* g) {
  switch (m->kind) {
    case M_STD_NORMAL: return lp_grad_std_normal(m, q, g);
    case M_BERNOULLI: return lp_grad_bernoulli(m, q, g);
    case M_EIGHT_SCHOOLS: return lp_grad_eight_schools(m, q, g);
    case M_LOGISTIC: return lp_grad_logistic(m, q, g);
    case M_IRT2PL: return lp_grad_irt(m, q, g);
    case M_POISSON: return lp_grad_poisson(m, q, g);
    case M_MOCK_FLAT:
      for (int i = 0; i < m->D; i++) g[i] = 0.0;
      return 0.0;
  }
  return NAN;
}

/* constrained output columns (model.write_array with tparams): returns count */
int oracle_num_outputs(const model_t* m) { return m->kind == M_EIGHT_SCHOOLS ? m->D + m->N : m->D; }

void oracle_write_array(const model_t* m, const double* q, double* out) {
  int D = m->D;
  switch (m->kind) {
    case M_BERNOULLI: out[0] = inv_logit(q[0]); break;
    case M_EIGHT_SCHOOLS: {
      int J = m->N;
      out[0] = q[0];
      out[1] = exp(q[1]);
      for (int j = 0; j < J; j++) out[2 + j] = q[2 + j];
      for (int j = 0; j < J; j++) out[2 + J + j] = out[0] + out[1] * q[2 + j];
      break;
    }
    case M_IRT2PL: {
      int I = m->N, J = m->K;
      for (int i = 0; i < D; i++) out[i] = q[i];
      out[0] = exp(q[0]);
      out[1 + I] = exp(q[1 + I]);
      for (int j = 0; j < J; j++) out[2 + I + j] = exp(q[2 + I + j]);
      out[3 + I + J] = exp(q[3 + I + J]);
      break;
    }
    default:
      for (int i = 0; i < D; i++) out[i] = q[i];
  }
}

/* ------------------------------------------------------------------ sampler ------------------------ */


typedef struct {
  int D;
  double *q, *p, *g;
  double V;
} point_t;

static point_t pt_alloc(int D) {
  point_t z;
  z.D = D;
  z.q = (double*)calloc(3 * (size_t)D, sizeof(double));
  z.p = z.q + D;
  z.g = z.p + D;
  z.V = 0;
  return z;
}
static void pt_free(point_t* z) { free(z->q); }
static void pt_copy(point_t* dst, const point_t* src) {
  memcpy(dst->q, src->q, 3 * (size_t)src->D * sizeof(double));
  dst->V = src->V;
}

typedef struct {
  const model_t* model;
  mixmax_t rng;
  int D;
  point_t z;
  double* inv_metric; /* diag_e: [D]; dense_e: [D][D] column-major like Eigen::MatrixXd, element (i,j) at j*D + i */
  int dense;          /* dense_e metric */
  double* chol;       /* dense_e: lower Cholesky factor L of inv_metric (L L^T = M^-1), column-major */
  double* tmpv;       /* dense_e: scratch vector [D] */
  double nom_epsilon, epsilon, epsilon_jitter;
  int max_depth;
  double max_deltaH;
  int depth, n_leapfrog, divergent;
  double energy;
  /* stepsize adaptation */
  double sa_counter, s_bar, x_bar, mu, delta, gamma, kappa, t0;
  /* windowed variance adaptation */
  unsigned num_warmup, adapt_init_buffer, adapt_term_buffer, adapt_base_window;
  unsigned adapt_window_counter, adapt_window_size, adapt_next_window;
  double wf_n, *wf_m, *wf_m2;
  long grad_evals;
} sampler_t;

static void update_potential_gradient(sampler_t* s, point_t* z) { /* base_hamiltonian.hpp:61-70 */
  double lp = oracle_lp_grad(s->model, z->q, z->g);
  s->grad_evals++;
  z->V = -lp;
  for (int i = 0; i < s->D; i++) z->g[i] = -z->g[i];
}
/* dtau_dp = M^-1 p.  diag_e_metric.hpp:36-38; dense_e_metric.hpp:38 (Eigen's column-major gemv accumulates each output
 * row sequentially over the columns -- the order used here) */
static void sharp(const sampler_t* s, const double* p, double* out) {
  const int D = s->D;
  if (!s->dense) {
    for (int i = 0; i < D; i++) out[i] = s->inv_metric[i] * p[i];
    return;
  }
  for (int i = 0; i < D; i++) {
    double acc = 0;
    for (int j = 0; j < D; j++) acc += s->inv_metric[(size_t)j * D + i] * p[j];
    out[i] = acc;
  }
}
/* lower Cholesky factor of the dense inverse metric (the reference re-factorises in every sample_p, dense_e_metric.hpp:55:
 * z.inv_e_metric_.llt(); the factor only changes when the metric does, so it is cached).  Only the lower triangle is
 * read, as Eigen::LLT<Lower> does.  Returns 0 when the matrix is not positive definite. */
static int dense_factor(sampler_t* s) {
  const int D = s->D;
  const double* A = s->inv_metric;
  double* L = s->chol;
  for (size_t k = 0; k < (size_t)D * D; k++) L[k] = 0;
  for (int j = 0; j < D; j++) {
    double d = A[(size_t)j * D + j];
    for (int k = 0; k < j; k++) d -= L[(size_t)k * D + j] * L[(size_t)k * D + j];
    if (!(d > 0)) return 0;
    const double ljj = sqrt(d);
    L[(size_t)j * D + j] = ljj;
    for (int i = j + 1; i < D; i++) {
      double v = A[(size_t)j * D + i];
      for (int k = 0; k < j; k++) v -= L[(size_t)k * D + i] * L[(size_t)k * D + j];
      L[(size_t)j * D + i] = v / ljj;
    }
  }
  return 1;
}
static double kinetic(const sampler_t* s, const point_t* z) { /* diag_e_metric.hpp:20-22; dense_e_metric.hpp:21-23 */
  double acc = 0;
  if (s->dense) {
    sharp(s, z->p, s->tmpv);
    for (int i = 0; i < s->D; i++) acc += z->p[i] * s->tmpv[i];
    return 0.5 * acc;
  }
  for (int i = 0; i < s->D; i++) acc += z->p[i] * (s->inv_metric[i] * z->p[i]);
  return 0.5 * acc;
}
static double hamiltonian(const sampler_t* s, const point_t* z) { return kinetic(s, z) + z->V; }
static void sample_p(sampler_t* s, point_t* z) { /* diag_e_metric.hpp:44-50; dense_e_metric.hpp:44-56 */
  const int D = s->D;
  if (s->dense) { /* u ~ N(0, I) in index order; p = L^-T u  (matrixU().solve(u): back substitution on L^T) */
    for (int i = 0; i < D; i++) s->tmpv[i] = oracle_normal(&s->rng);
    const double* L = s->chol;
    for (int i = D - 1; i >= 0; i--) {
      double acc = 0;
      for (int k = i + 1; k < D; k++) acc += L[(size_t)i * D + k] * z->p[k];
      z->p[i] = (s->tmpv[i] - acc) / L[(size_t)i * D + i];
    }
    return;
  }
  for (int i = 0; i < D; i++) z->p[i] = oracle_normal(&s->rng) / sqrt(s->inv_metric[i]);
}
static void evolve(sampler_t* s, point_t* z, double eps) { /* base_leapfrog.hpp:17-22 + expl_leapfrog.hpp:16-32 */
  double h = 0.5 * eps;
  for (int i = 0; i < s->D; i++) z->p[i] -= h * z->g[i];
  if (s->dense) {
    sharp(s, z->p, s->tmpv);
    for (int i = 0; i < s->D; i++) z->q[i] += eps * s->tmpv[i];
  } else {
    for (int i = 0; i < s->D; i++) z->q[i] += eps * (s->inv_metric[i] * z->p[i]);
  }
  update_potential_gradient(s, z);
  for (int i = 0; i < s->D; i++) z->p[i] -= h * z->g[i];
}
static double log_sum_exp2(double a, double b) { /* log_sum_exp.hpp:51-64 */
  if (a == -INFINITY) return b;
  if (a == INFINITY && b == INFINITY) return INFINITY;
  if (a > b) return a + log1p_exp(b - a);
  return b + log1p_exp(a - b);
}
static double dot(const double* a, const double* b, int n) {
  double s = 0;
  for (int i = 0; i < n; i++) s += a[i] * b[i];
  return s;
}
static int criterion(const double* p_sharp_minus, const double* p_sharp_plus, const double* rho, int n) {
  return dot(p_sharp_plus, rho, n) > 0 && dot(p_sharp_minus, rho, n) > 0;
}

static int init_stepsize(sampler_t* s) { /* base_hmc.hpp:93-161; returns 0 ok, 1 improper, 2 no small step */
  if (s->D == 0) {
    s->nom_epsilon = NAN;
    return 0;
  }
  if (s->nom_epsilon == 0 || s->nom_epsilon > 1e7 || isnan(s->nom_epsilon)) return 0;
  point_t z_init = pt_alloc(s->D);
  pt_copy(&z_init, &s->z);
  sample_p(s, &s->z);
  update_potential_gradient(s, &s->z);
  double H0 = hamiltonian(s, &s->z);
  evolve(s, &s->z, s->nom_epsilon);
  double h = hamiltonian(s, &s->z);
  if (isnan(h)) h = INFINITY;
  double delta_H = H0 - h;
  int direction = delta_H > log(0.8) ? 1 : -1;
  int rc = 0;
  for (;;) {
    pt_copy(&s->z, &z_init);
    sample_p(s, &s->z);
    update_potential_gradient(s, &s->z);
    H0 = hamiltonian(s, &s->z);
    evolve(s, &s->z, s->nom_epsilon);
    h = hamiltonian(s, &s->z);
    if (isnan(h)) h = INFINITY;
    delta_H = H0 - h;
    if (direction == 1 && !(delta_H > log(0.8))) break;
    else if (direction == -1 && !(delta_H < log(0.8))) break;
    else s->nom_epsilon = direction == 1 ? 2.0 * s->nom_epsilon : 0.5 * s->nom_epsilon;
    if (s->nom_epsilon > 1e7) { rc = 1; break; }
    if (s->nom_epsilon == 0) { rc = 2; break; }
  }
  pt_copy(&s->z, &z_init);
  pt_free(&z_init);
  return rc;
}

/* base_nuts.hpp:247-352 -- recursive, exactly the reference's structure */
static int build_tree(sampler_t* s, int depth, point_t* z_propose, double* p_sharp_beg, double* p_sharp_end, double* rho,
                      double* p_beg, double* p_end, double H0, double sign, int* n_leapfrog, double* log_sum_weight,
                      double* sum_metro_prob) {
  int D = s->D;
  if (depth == 0) {
    evolve(s, &s->z, sign * s->epsilon);
    ++*n_leapfrog;
    double h = hamiltonian(s, &s->z);
    if (isnan(h)) h = INFINITY;
    if ((h - H0) > s->max_deltaH) s->divergent = 1;
    *log_sum_weight = log_sum_exp2(*log_sum_weight, H0 - h);
    if (H0 - h > 0) *sum_metro_prob += 1;
    else *sum_metro_prob += exp(H0 - h);
    pt_copy(z_propose, &s->z);
    sharp(s, s->z.p, p_sharp_beg);
    for (int i = 0; i < D; i++) {
      p_sharp_end[i] = p_sharp_beg[i];
      rho[i] += s->z.p[i];
      p_beg[i] = s->z.p[i];
      p_end[i] = p_beg[i];
    }
    return !s->divergent;
  }
  double* buf = (double*)calloc(7 * (size_t)D, sizeof(double));
  double *p_init_end = buf, *p_sharp_init_end = buf + D, *rho_init = buf + 2 * D;
  double *p_final_beg = buf + 3 * D, *p_sharp_final_beg = buf + 4 * D, *rho_final = buf + 5 * D, *rho_sub = buf + 6 * D;
  double lsw_init = -INFINITY;
  int ok = build_tree(s, depth - 1, z_propose, p_sharp_beg, p_sharp_init_end, rho_init, p_beg, p_init_end, H0, sign,
                      n_leapfrog, &lsw_init, sum_metro_prob);
  if (!ok) {
    free(buf);
    return 0;
  }
  point_t z_propose_final = pt_alloc(D);
  pt_copy(&z_propose_final, &s->z);
  double lsw_final = -INFINITY;
  ok = build_tree(s, depth - 1, &z_propose_final, p_sharp_final_beg, p_sharp_end, rho_final, p_final_beg, p_end, H0, sign,
                  n_leapfrog, &lsw_final, sum_metro_prob);
  if (!ok) {
    pt_free(&z_propose_final);
    free(buf);
    return 0;
  }
  double lsw_subtree = log_sum_exp2(lsw_init, lsw_final);
  *log_sum_weight = log_sum_exp2(*log_sum_weight, lsw_subtree);
  if (lsw_final > lsw_subtree) {
    pt_copy(z_propose, &z_propose_final);
  } else {
    double accept_prob = exp(lsw_final - lsw_subtree);
    if (oracle_uniform01(&s->rng) < accept_prob) pt_copy(z_propose, &z_propose_final);
  }
  for (int i = 0; i < D; i++) {
    rho_sub[i] = rho_init[i] + rho_final[i];
    rho[i] += rho_sub[i];
  }
  int persist = criterion(p_sharp_beg, p_sharp_end, rho_sub, D);
  for (int i = 0; i < D; i++) rho_sub[i] = rho_init[i] + p_final_beg[i];
  persist &= criterion(p_sharp_beg, p_sharp_final_beg, rho_sub, D);
  for (int i = 0; i < D; i++) rho_sub[i] = rho_final[i] + p_init_end[i];
  persist &= criterion(p_sharp_init_end, p_sharp_end, rho_sub, D);
  pt_free(&z_propose_final);
  free(buf);
  return persist;
}

/* base_nuts.hpp:78-204.  q_in/out = s->z.q ; returns accept_stat, sets lp */
static double transition_from_momentum(sampler_t* s, double* lp_out);
static double transition(sampler_t* s, double* lp_out) {
  s->epsilon = s->nom_epsilon; /* base_hmc.hpp:195-200 */
  if (s->epsilon_jitter) s->epsilon *= 1.0 + s->epsilon_jitter * (2.0 * oracle_uniform01(&s->rng) - 1.0);
  sample_p(s, &s->z);
  return transition_from_momentum(s, lp_out);
}
/* base_nuts.hpp:88-204: everything after `this->hamiltonian_.sample_p(this->z_, this->rand_int_)` (separate so that the
   reference's mock-Hamiltonian KATs, whose sample_p is a no-op, can enter here with a given momentum) */
static double transition_from_momentum(sampler_t* s, double* lp_out) {
  int D = s->D;
  update_potential_gradient(s, &s->z);
  point_t z_fwd = pt_alloc(D), z_bck = pt_alloc(D), z_sample = pt_alloc(D), z_propose = pt_alloc(D);
  pt_copy(&z_fwd, &s->z);
  pt_copy(&z_bck, &s->z);
  pt_copy(&z_sample, &s->z);
  pt_copy(&z_propose, &s->z);
  double* buf = (double*)calloc(12 * (size_t)D, sizeof(double));
  double *p_fwd_fwd = buf, *p_sharp_fwd_fwd = buf + D, *p_fwd_bck = buf + 2 * D, *p_sharp_fwd_bck = buf + 3 * D;
  double *p_bck_fwd = buf + 4 * D, *p_sharp_bck_fwd = buf + 5 * D, *p_bck_bck = buf + 6 * D, *p_sharp_bck_bck = buf + 7 * D;
  double *rho = buf + 8 * D, *rho_fwd = buf + 9 * D, *rho_bck = buf + 10 * D, *rho_ext = buf + 11 * D;
  sharp(s, s->z.p, p_sharp_fwd_fwd);
  for (int i = 0; i < D; i++) {
    double ps = p_sharp_fwd_fwd[i];
    p_fwd_fwd[i] = p_fwd_bck[i] = p_bck_fwd[i] = p_bck_bck[i] = s->z.p[i];
    p_sharp_fwd_fwd[i] = p_sharp_fwd_bck[i] = p_sharp_bck_fwd[i] = p_sharp_bck_bck[i] = ps;
    rho[i] = s->z.p[i];
  }
  double log_sum_weight = 0;
  double H0 = hamiltonian(s, &s->z);
  int n_leapfrog = 0;
  double sum_metro_prob = 0;
  s->depth = 0;
  s->divergent = 0;
  while (s->depth < s->max_depth) {
    for (int i = 0; i < D; i++) rho_fwd[i] = rho_bck[i] = 0;
    int valid_subtree = 0;
    double lsw_subtree = -INFINITY;
    if (oracle_uniform01(&s->rng) > 0.5) {
      pt_copy(&s->z, &z_fwd);
      for (int i = 0; i < D; i++) {
        rho_bck[i] = rho[i];
        p_bck_fwd[i] = p_fwd_fwd[i];
        p_sharp_bck_fwd[i] = p_sharp_fwd_fwd[i];
      }
      valid_subtree = build_tree(s, s->depth, &z_propose, p_sharp_fwd_bck, p_sharp_fwd_fwd, rho_fwd, p_fwd_bck, p_fwd_fwd,
                                 H0, 1, &n_leapfrog, &lsw_subtree, &sum_metro_prob);
      pt_copy(&z_fwd, &s->z);
    } else {
      pt_copy(&s->z, &z_bck);
      for (int i = 0; i < D; i++) {
        rho_fwd[i] = rho[i];
        p_fwd_bck[i] = p_bck_bck[i];
        p_sharp_fwd_bck[i] = p_sharp_bck_bck[i];
      }
      valid_subtree = build_tree(s, s->depth, &z_propose, p_sharp_bck_fwd, p_sharp_bck_bck, rho_bck, p_bck_fwd, p_bck_bck,
                                 H0, -1, &n_leapfrog, &lsw_subtree, &sum_metro_prob);
      pt_copy(&z_bck, &s->z);
    }
    if (!valid_subtree) break;
    ++s->depth;
    if (lsw_subtree > log_sum_weight) {
      pt_copy(&z_sample, &z_propose);
    } else {
      double accept_prob = exp(lsw_subtree - log_sum_weight);
      if (oracle_uniform01(&s->rng) < accept_prob) pt_copy(&z_sample, &z_propose);
    }
    log_sum_weight = log_sum_exp2(log_sum_weight, lsw_subtree);
    for (int i = 0; i < D; i++) rho[i] = rho_bck[i] + rho_fwd[i];
    int persist = criterion(p_sharp_bck_bck, p_sharp_fwd_fwd, rho, D);
    for (int i = 0; i < D; i++) rho_ext[i] = rho_bck[i] + p_fwd_bck[i];
    persist &= criterion(p_sharp_bck_bck, p_sharp_fwd_bck, rho_ext, D);
    for (int i = 0; i < D; i++) rho_ext[i] = rho_fwd[i] + p_bck_fwd[i];
    persist &= criterion(p_sharp_bck_fwd, p_sharp_fwd_fwd, rho_ext, D);
    if (!persist) break;
  }
  s->n_leapfrog = n_leapfrog;
  double accept_prob = sum_metro_prob / (double)n_leapfrog;
  pt_copy(&s->z, &z_sample);
  s->energy = hamiltonian(s, &s->z);
  *lp_out = -s->z.V;
  pt_free(&z_fwd);
  pt_free(&z_bck);
  pt_free(&z_sample);
  pt_free(&z_propose);
  free(buf);
  return accept_prob;
}

/* base_static_hmc::transition, S/mcmc/hmc/static/base_static_hmc.hpp:41-70 */
static double transition_static(sampler_t* s, int L, double* lp_out) {
  int D = s->D;
  s->epsilon = s->nom_epsilon;
  if (s->epsilon_jitter) s->epsilon *= 1.0 + s->epsilon_jitter * (2.0 * oracle_uniform01(&s->rng) - 1.0);
  sample_p(s, &s->z);
  update_potential_gradient(s, &s->z);
  point_t z_init = pt_alloc(D);
  pt_copy(&z_init, &s->z);
  double H0 = hamiltonian(s, &s->z);
  for (int i = 0; i < L; i++) evolve(s, &s->z, s->epsilon);
  double h = hamiltonian(s, &s->z);
  if (isnan(h)) h = INFINITY;
  double accept_prob = exp(H0 - h);
  if (accept_prob < 1 && oracle_uniform01(&s->rng) > accept_prob) pt_copy(&s->z, &z_init);
  accept_prob = accept_prob > 1 ? 1 : accept_prob;
  s->energy = hamiltonian(s, &s->z);
  s->depth = L;
  s->n_leapfrog = 1;
  s->divergent = 0;
  *lp_out = -s->z.V;
  pt_free(&z_init);
  return accept_prob;
}
static int static_L(double T, double nom_eps) { /* base_static_hmc::update_L_ */
  int L = (int)(T / nom_eps);
  return L < 1 ? 1 : L;
}

static void sa_restart(sampler_t* s) { s->sa_counter = 0; s->s_bar = 0; s->x_bar = 0; }
static void learn_stepsize(sampler_t* s, double adapt_stat) { /* stepsize_adaptation.hpp:55-71 */
  ++s->sa_counter;
  adapt_stat = adapt_stat > 1 ? 1 : adapt_stat;
  const double eta = 1.0 / (s->sa_counter + s->t0);
  s->s_bar = (1.0 - eta) * s->s_bar + eta * (s->delta - adapt_stat);
  const double x = s->mu - s->s_bar * sqrt(s->sa_counter) / s->gamma;
  const double x_eta = pow(s->sa_counter, -s->kappa);
  s->x_bar = (1.0 - x_eta) * s->x_bar + x_eta * x;
  s->nom_epsilon = exp(x);
}
static void win_restart(sampler_t* s) {
  s->adapt_window_counter = 0;
  s->adapt_window_size = s->adapt_base_window;
  s->adapt_next_window = s->adapt_init_buffer + s->adapt_window_size - 1;
}
static void set_window_params(sampler_t* s, unsigned num_warmup, unsigned init_buffer, unsigned term_buffer,
                              unsigned base_window) { /* windowed_adaptation.hpp:35-79 */
  if (num_warmup < 20) return;
  if (init_buffer + base_window + term_buffer > num_warmup) {
    s->num_warmup = num_warmup;
    s->adapt_init_buffer = 0.15 * num_warmup;
    s->adapt_term_buffer = 0.10 * num_warmup;
    s->adapt_base_window = num_warmup - (s->adapt_init_buffer + s->adapt_term_buffer);
    return; /* NB: the reference returns here without restart() */
  }
  s->num_warmup = num_warmup;
  s->adapt_init_buffer = init_buffer;
  s->adapt_term_buffer = term_buffer;
  s->adapt_base_window = base_window;
  win_restart(s);
}
static int learn_variance(sampler_t* s) { /* var_adaptation.hpp:17-46 */
  int D = s->D;
  const double* q = s->z.q;
  int in_window = (s->adapt_window_counter >= s->adapt_init_buffer) &&
                  (s->adapt_window_counter < s->num_warmup - s->adapt_term_buffer) &&
                  (s->adapt_window_counter != s->num_warmup);
  if (in_window && s->dense) { /* welford_covar_estimator::add_sample: m2 += (q - m_new) delta^T */
    ++s->wf_n;
    double* delta = s->tmpv;
    for (int i = 0; i < D; i++) {
      delta[i] = q[i] - s->wf_m[i];
      s->wf_m[i] += delta[i] / s->wf_n;
    }
    for (int j = 0; j < D; j++)
      for (int i = 0; i < D; i++) s->wf_m2[(size_t)j * D + i] += (q[i] - s->wf_m[i]) * delta[j];
  } else if (in_window) {
    ++s->wf_n;
    for (int i = 0; i < D; i++) {
      double delta = q[i] - s->wf_m[i];
      s->wf_m[i] += delta / s->wf_n;
      s->wf_m2[i] += delta * (q[i] - s->wf_m[i]);
    }
  }
  int end_window = (s->adapt_window_counter == s->adapt_next_window) && (s->adapt_window_counter != s->num_warmup);
  if (end_window) {
    /* compute_next_window, windowed_adaptation.hpp:91-110 */
    if (!(s->adapt_next_window == s->num_warmup - s->adapt_term_buffer - 1)) {
      s->adapt_window_size *= 2;
      s->adapt_next_window = s->adapt_window_counter + s->adapt_window_size;
      if (!(s->adapt_next_window == s->num_warmup - s->adapt_term_buffer - 1)) {
        unsigned next_window_boundary = s->adapt_next_window + 2 * s->adapt_window_size;
        if (next_window_boundary >= s->num_warmup - s->adapt_term_buffer)
          s->adapt_next_window = s->num_warmup - s->adapt_term_buffer - 1;
      }
    }
    double n = s->wf_n;
    if (s->dense) { /* covar_adaptation.hpp:24-31 */
      for (int j = 0; j < D; j++)
        for (int i = 0; i < D; i++) {
          const size_t e = (size_t)j * D + i;
          double cov = s->inv_metric[e];
          if (n > 1) cov = s->wf_m2[e] / (n - 1.0);
          s->inv_metric[e] = (n / (n + 5.0)) * cov + 1e-3 * (5.0 / (n + 5.0)) * (i == j ? 1.0 : 0.0);
          s->wf_m2[e] = 0;
        }
      for (int i = 0; i < D; i++) s->wf_m[i] = 0;
      s->wf_n = 0;
      ++s->adapt_window_counter;
      return dense_factor(s) ? 1 : -1;
    }
    for (int i = 0; i < D; i++) {
      double var = s->inv_metric[i];
      if (n > 1) var = s->wf_m2[i] / (n - 1.0);
      s->inv_metric[i] = (n / (n + 5.0)) * var + 1e-3 * (5.0 / (n + 5.0)) * 1.0;
    }
    s->wf_n = 0;
    for (int i = 0; i < D; i++) s->wf_m[i] = s->wf_m2[i] = 0;
    ++s->adapt_window_counter;
    return 1;
  }
  ++s->adapt_window_counter;
  return 0;
}

/*
 * Runs one chain exactly as services::sample::hmc_nuts_diag_e_adapt would (hmc_nuts_diag_e_adapt.hpp:58-112,
 * run_adaptive_sampler.hpp:47-103).  q0 != NULL: user-specified unconstrained inits (one try; the D uniforms of the unused
 * random context are still drawn when init_radius > 0).
 * draws: [(save_warmup? W:0 + S)/thin rows][7 + D] (7 + 3 D with c->diag: then q is followed by the sample's p and g) =
 *        lp__, accept_stat__, stepsize__, treedepth__, n_leapfrog__, divergent__, energy__, q (unconstrained).
 * adapt_out: [1 + D] = final step size, inverse metric ([1 + D*D], row-major, for dense_e).  inv_metric0 likewise [D] or
 * [D][D].  Returns 0 ok, 78 init failure, 70 stepsize failure.
 */
int oracle_run_chain(const model_t* model, const oracle_config_t* c, const double* q0, const double* inv_metric0,
                     double* draws, double* adapt_out, double* q_init_out, long* grad_evals_out) {
  int D = model->D;
  sampler_t s;
  memset(&s, 0, sizeof s);
  s.model = model;
  s.D = D;
  s.z = pt_alloc(D);
  s.dense = c->dense_e;
  const size_t msz = s.dense ? (size_t)D * D : (size_t)D;
  s.inv_metric = (double*)malloc(msz * sizeof(double));
  s.wf_m = (double*)calloc(D, sizeof(double));
  s.wf_m2 = (double*)calloc(msz, sizeof(double));
  s.chol = s.dense ? (double*)calloc(msz, sizeof(double)) : NULL;
  s.tmpv = (double*)calloc(D, sizeof(double));
  double* gtmp = (double*)malloc(D * sizeof(double));
  int rc = 0;
  oracle_rng_seed(&s.rng, c->seed, c->chain_id);
  /* ---- initialize ---- */
  int ok = 0;
  /* a NaN in q0 stands for "not given" (partial user inits): the reference keeps the random context's constrained value for such
   * a parameter and transform_inits maps it back (initialize.hpp:93-101) -- the round trip of the uniform draw through the
   * parameter's transform, coded in the NaN payload: 0 none, 1 lower bound 0, 2 unit interval */
  int partial = 0;
  if (q0)
    for (int i = 0; i < D; i++) partial |= (q0[i] != q0[i]);
  int max_tries = ((q0 && !partial) || c->init_radius == 0.0) ? 1 : 100;
  for (int t = 0; t < max_tries && !ok; t++) {
    if (q0) {
      /* initialize.hpp:93-101: the random_var_context is constructed -- and draws its D uniforms (random_var_context.hpp:43-66)
       * -- on every try, also when every value then comes from the user's context through the chained_var_context */
      for (int i = 0; i < D; i++) {
        double u = c->init_radius != 0.0 ? oracle_uniform_real(&s.rng, -c->init_radius, c->init_radius) : 0.0;
        if (q0[i] == q0[i]) {
          s.z.q[i] = q0[i];
        } else {
          uint64_t bits;
          memcpy(&bits, &q0[i], sizeof bits);
          int code = (int)(bits & 3u);
          if (code == 1) s.z.q[i] = log(exp(u));
          else if (code == 2) { double th = inv_logit(u); s.z.q[i] = log(th / (1.0 - th)); }
          else s.z.q[i] = u;
        }
      }
    }
    else if (c->init_radius == 0.0) memset(s.z.q, 0, D * sizeof(double));
    else for (int i = 0; i < D; i++) s.z.q[i] = oracle_uniform_real(&s.rng, -c->init_radius, c->init_radius);
    double lp = oracle_lp_grad(model, s.z.q, gtmp);
    if (!isfinite(lp)) continue;
    double gs = 0;
    for (int i = 0; i < D; i++) gs += gtmp[i];
    if (isfinite(gs)) ok = 1;
  }
  if (!ok) { rc = 78; goto done; }
  if (q_init_out) memcpy(q_init_out, s.z.q, D * sizeof(double));
  /* ---- sampler configuration, hmc_nuts_diag_e_adapt.hpp:84-98 ---- */
  if (s.dense) {
    for (int i = 0; i < D; i++)
      for (int j = 0; j < D; j++) s.inv_metric[(size_t)j * D + i] = inv_metric0 ? inv_metric0[(size_t)i * D + j] : (i == j ? 1.0 : 0.0);
    if (!dense_factor(&s)) { rc = 78; goto done; }
  } else {
    for (int i = 0; i < D; i++) s.inv_metric[i] = inv_metric0 ? inv_metric0[i] : 1.0;
  }
  s.nom_epsilon = 0.1;
  if (c->stepsize > 0) s.nom_epsilon = c->stepsize;
  s.epsilon = s.nom_epsilon;
  s.epsilon_jitter = (c->stepsize_jitter > 0 && c->stepsize_jitter < 1) ? c->stepsize_jitter : 0.0;
  s.max_depth = c->max_depth > 0 ? c->max_depth : 5;
  s.max_deltaH = 1000;
  s.mu = 0.5; s.delta = 0.5; s.gamma = 0.05; s.kappa = 0.75; s.t0 = 10;
  sa_restart(&s);
  s.mu = log(10 * c->stepsize);
  if (c->delta > 0 && c->delta < 1) s.delta = c->delta;
  if (c->gamma > 0) s.gamma = c->gamma;
  if (c->kappa > 0) s.kappa = c->kappa;
  if (c->t0 > 0) s.t0 = c->t0;
  s.adapt_window_counter = 0; s.adapt_window_size = 0; s.adapt_next_window = 0u - 1u;
  if (c->adapt_engaged && !c->unit_e) set_window_params(&s, c->num_warmup, c->init_buffer, c->term_buffer, c->window);
  /* ---- run_adaptive_sampler ---- */
  int adapt_flag = c->adapt_engaged;
  update_potential_gradient(&s, &s.z); /* z.V, z.g at q0 (init_stepsize / transition re-evaluate them anyway) */
  if (adapt_flag && init_stepsize(&s)) { rc = 70; goto done; } /* hmc_nuts_diag_e (no adapt) uses run_sampler: no init_stepsize */
  int row = 0, ncol = 7 + D * (c->diag ? 3 : 1);
  /* static HMC: set_nominal_stepsize_and_T(stepsize, int_time) fixes L from the stepsize ARGUMENT; init_stepsize above changed
   * nom_epsilon without touching L (base_hmc::init_stepsize writes nom_epsilon_ directly) */
  int L = c->static_hmc ? static_L(c->int_time, c->stepsize > 0 ? c->stepsize : 0.1) : 0;
  int total = c->num_warmup + c->num_samples;
  for (int it = 0; it < total; it++) {
    int warm = it < c->num_warmup;
    double lp;
    double accept = c->static_hmc ? transition_static(&s, L, &lp) : transition(&s, &lp);
    double eps_used = s.epsilon;
    if (adapt_flag && warm) {
      learn_stepsize(&s, accept);
      if (c->static_hmc) L = static_L(c->int_time, s.nom_epsilon); /* adapt_*_static_hmc::transition: update_L_() */
      int upd = learn_variance(&s);
      if (upd < 0) { rc = 70; goto done; }
      if (upd) {
        if (init_stepsize(&s)) { rc = 70; goto done; }
        if (c->static_hmc) L = static_L(c->int_time, s.nom_epsilon);
        s.mu = log(10 * s.nom_epsilon);
        sa_restart(&s);
      }
    }
    int m = warm ? it : it - c->num_warmup;
    int save = (warm ? c->save_warmup : 1) && (m % c->num_thin == 0);
    if (save && draws) {
      double* r = draws + (size_t)row * ncol;
      r[0] = lp; r[1] = accept; r[2] = eps_used; r[3] = s.depth; r[4] = s.n_leapfrog; r[5] = s.divergent; r[6] = s.energy;
      memcpy(r + 7, s.z.q, D * sizeof(double));
      if (c->diag) {
        memcpy(r + 7 + D, s.z.p, D * sizeof(double));
        memcpy(r + 7 + 2 * D, s.z.g, D * sizeof(double));
      }
      row++;
    }
    if (it == c->num_warmup - 1 && adapt_flag) s.nom_epsilon = exp(s.x_bar); /* disengage_adaptation */
  }
  if (adapt_out) {
    adapt_out[0] = s.nom_epsilon;
    if (s.dense) {
      for (int i = 0; i < D; i++)
        for (int j = 0; j < D; j++) adapt_out[1 + (size_t)i * D + j] = s.inv_metric[(size_t)j * D + i];
    } else {
      memcpy(adapt_out + 1, s.inv_metric, D * sizeof(double));
    }
  }
done:
  if (grad_evals_out) *grad_evals_out = s.grad_evals;
  pt_free(&s.z);
  free(s.inv_metric); free(s.wf_m); free(s.wf_m2); free(s.chol); free(s.tmpv); free(gtmp);
  return rc;
}

/* ---- state-injection entry points for the reference's adaptation KATs (ST/unit/mcmc/stepsize_adaptation_test.cpp:99-143,
 *      var_adaptation_test.cpp) ---- */
void oracle_learn_stepsize_kat(double counter, double s_bar, double x_bar, double mu, double delta, double gamma, double kappa,
                               double t0, double adapt_stat, double* out /* eps, counter, s_bar, x_bar */) {
  sampler_t s;
  memset(&s, 0, sizeof s);
  s.sa_counter = counter; s.s_bar = s_bar; s.x_bar = x_bar; s.mu = mu; s.delta = delta; s.gamma = gamma; s.kappa = kappa; s.t0 = t0;
  learn_stepsize(&s, adapt_stat);
  out[0] = s.nom_epsilon; out[1] = s.sa_counter; out[2] = s.s_bar; out[3] = s.x_bar;
}

/* feeds n_learn copies of q through learn_variance with set_window_params(num_warmup, init_buffer, term_buffer, window);
 * returns how many times the metric was updated; var_out = final inverse metric */
int oracle_learn_variance_kat(int D, const double* q, int n_learn, unsigned num_warmup, unsigned init_buffer, unsigned term_buffer,
                              unsigned window, double* var_out) {
  sampler_t s;
  memset(&s, 0, sizeof s);
  s.D = D;
  s.z = pt_alloc(D);
  memcpy(s.z.q, q, D * sizeof(double));
  s.inv_metric = (double*)calloc(D, sizeof(double));
  s.wf_m = (double*)calloc(D, sizeof(double));
  s.wf_m2 = (double*)calloc(D, sizeof(double));
  s.adapt_window_counter = 0; s.adapt_window_size = 0; s.adapt_next_window = 0u - 1u;
  set_window_params(&s, num_warmup, init_buffer, term_buffer, window);
  int updates = 0;
  for (int i = 0; i < n_learn; i++) updates += learn_variance(&s);
  memcpy(var_out, s.inv_metric, D * sizeof(double));
  pt_free(&s.z); free(s.inv_metric); free(s.wf_m); free(s.wf_m2);
  return updates;
}

/* one expl_leapfrog::evolve on z = (q, p, g, V) of `model` with step eps and a diagonal inverse metric
 * (ST/unit/mcmc/hmc/integrators/expl_leapfrog_test.cpp:156-370: evolve_1 .. evolve_7); out = q[D], p[D], g[D], V */
void oracle_evolve_kat(const model_t* model, const double* q, const double* p, const double* inv_metric, double eps, double* out) {
  sampler_t s;
  memset(&s, 0, sizeof s);
  const int D = model->D;
  s.model = model;
  s.D = D;
  s.z = pt_alloc(D);
  s.inv_metric = (double*)malloc(D * sizeof(double));
  for (int i = 0; i < D; i++) {
    s.inv_metric[i] = inv_metric ? inv_metric[i] : 1.0;
    s.z.q[i] = q[i];
    s.z.p[i] = p[i];
  }
  update_potential_gradient(&s, &s.z);
  evolve(&s, &s.z, eps);
  memcpy(out, s.z.q, D * sizeof(double));
  memcpy(out + D, s.z.p, D * sizeof(double));
  memcpy(out + 2 * D, s.z.g, D * sizeof(double));
  out[3 * D] = s.z.V;
  pt_free(&s.z);
  free(s.inv_metric);
}

/* ---- ST/unit/mcmc/hmc/nuts/base_nuts_test.cpp restated.  The reference runs these on mock classes (mock_hmc.hpp): log_prob == 0,
   sample_p a no-op, evolve: q += eps p, H == 0 everywhere, U-turn criterion overridden to `true`.  With the flat model, a unit
   metric and a fixed momentum the real classes behave identically where the tests look: the leapfrog leaves p unchanged and moves q
   by eps p, H is the same at every point, and rho . p_sharp > 0 always.  Where the mock's p_sharp (= q) differs from the real one
   (= p) the expectation is not restated. */
static void kat_sampler(sampler_t* s, const model_t* model, uint32_t seed, uint32_t chain, const double* q0, const double* p0, double eps,
                        int max_depth) {
  memset(s, 0, sizeof *s);
  s->model = model;
  s->D = model->D;
  s->z = pt_alloc(model->D);
  s->inv_metric = (double*)malloc(model->D * sizeof(double));
  for (int i = 0; i < model->D; i++) {
    s->inv_metric[i] = 1.0;
    s->z.q[i] = q0[i];
    s->z.p[i] = p0[i];
  }
  s->nom_epsilon = s->epsilon = eps;
  s->max_depth = max_depth;
  s->max_deltaH = 1000;
  oracle_rng_seed(&s->rng, seed, chain);
}
/* TEST(McmcNutsBaseNuts, transition), base_nuts_test.cpp:335-372: a transition that starts from a GIVEN momentum.
   out: accept_stat, lp, depth, n_leapfrog, divergent, q[D] */
void oracle_nuts_transition_kat(const model_t* model, uint32_t seed, uint32_t chain, const double* q0, const double* p0, double eps,
                                int max_depth, double* out) {
  sampler_t s;
  kat_sampler(&s, model, seed, chain, q0, p0, eps, max_depth);
  double lp = 0;
  out[0] = transition_from_momentum(&s, &lp);
  out[1] = lp;
  out[2] = s.depth;
  out[3] = s.n_leapfrog;
  out[4] = s.divergent;
  memcpy(out + 5, s.z.q, s.D * sizeof(double));
  pt_free(&s.z);
  free(s.inv_metric);
}
/* TEST(McmcNutsBaseNuts, build_tree_test), base_nuts_test.cpp:139-195, and TEST(McmcUnitENuts, build_tree_test),
   unit_e_nuts_test.cpp:10-95: build_tree(depth, ..., H0, sign, ...) from (q0, p0) with rho = p0; H0 is taken as given
   (h0_relative == 0) or as an offset from H(q0, p0).
   out: valid, n_leapfrog, log_sum_weight, sum_metro_prob, then rho, p_begin, p_sharp_begin, p_end, p_sharp_end, z.q, z.p (D each) */
void oracle_nuts_build_tree_kat(const model_t* model, uint32_t seed, uint32_t chain, const double* q0, const double* p0, double eps, int depth,
                                double sign, double H0, int h0_relative, double* out) {
  sampler_t s;
  kat_sampler(&s, model, seed, chain, q0, p0, eps, 10);
  update_potential_gradient(&s, &s.z);
  const int D = s.D;
  point_t z_propose = pt_alloc(D);
  double* buf = (double*)calloc(5 * (size_t)D, sizeof(double));
  double *p_sharp_begin = buf, *p_sharp_end = buf + D, *rho = buf + 2 * D, *p_begin = buf + 3 * D, *p_end = buf + 4 * D;
  for (int i = 0; i < D; i++) rho[i] = s.z.p[i];
  double lsw = -INFINITY, sum_metro = 0;
  int n_leapfrog = 0;
  if (h0_relative) H0 += hamiltonian(&s, &s.z);
  out[0] = build_tree(&s, depth, &z_propose, p_sharp_begin, p_sharp_end, rho, p_begin, p_end, H0, sign, &n_leapfrog, &lsw, &sum_metro);
  out[1] = n_leapfrog;
  out[2] = lsw;
  out[3] = sum_metro;
  const double* vecs[7] = {rho, p_begin, p_sharp_begin, p_end, p_sharp_end, s.z.q, s.z.p};
  for (int k = 0; k < 7; k++) memcpy(out + 4 + (size_t)k * D, vecs[k], D * sizeof(double));
  pt_free(&z_propose);
  pt_free(&s.z);
  free(buf);
  free(s.inv_metric);
}
/* TEST(McmcNutsDerivedNuts, compute_criterion_{unit,diag,dense}_e), ST/unit/mcmc/hmc/nuts/derived_nuts_test.cpp:14-132 */
int oracle_criterion_kat(const double* p_sharp_minus, const double* p_sharp_plus, const double* rho, int n) {
  return criterion(p_sharp_minus, p_sharp_plus, rho, n);
}
/* TEST(McmcCovarAdaptation, learn_covariance), ST/unit/mcmc/covar_adaptation_test.cpp:5-31: n_learn samples q fed to the dense
   (covariance) window logic; covar_out [D][D] */
int oracle_learn_covariance_kat(int D, const double* q, int n_learn, unsigned num_warmup, unsigned init_buffer, unsigned term_buffer,
                                unsigned window, double* covar_out) {
  sampler_t s;
  memset(&s, 0, sizeof s);
  s.D = D;
  s.dense = 1;
  s.z = pt_alloc(D);
  memcpy(s.z.q, q, D * sizeof(double));
  s.inv_metric = (double*)calloc((size_t)D * D, sizeof(double));
  s.chol = (double*)calloc((size_t)D * D, sizeof(double));
  s.tmpv = (double*)calloc(D, sizeof(double));
  s.wf_m = (double*)calloc(D, sizeof(double));
  s.wf_m2 = (double*)calloc((size_t)D * D, sizeof(double));
  s.adapt_window_counter = 0; s.adapt_window_size = 0; s.adapt_next_window = 0u - 1u;
  set_window_params(&s, num_warmup, init_buffer, term_buffer, window);
  int updates = 0;
  for (int i = 0; i < n_learn; i++) updates += learn_variance(&s) != 0;
  memcpy(covar_out, s.inv_metric, (size_t)D * D * sizeof(double));
  pt_free(&s.z); free(s.inv_metric); free(s.chol); free(s.tmpv); free(s.wf_m); free(s.wf_m2);
  return updates;
}
