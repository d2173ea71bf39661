// Device-side model family: closed-form log density (propto = true, jacobian = true -- exactly what the sampler
// sees through stan::model::log_prob_propto_jacobian / stan::model::gradient, S/model/gradient.hpp:21-35) and its
// gradient on the unconstrained scale.  There is no AD tape on the device: each density is hand-derived, and
// checked against the reference's own log_prob_grad to <= 1e-10 relative (tests/test_lp_grad*.py).
//
// Small models (kinds 0-2) are warp-cooperative device functions that are inlined into the fused chain kernel.
// Data-heavy models (kinds 3-5) have their own cross-chain kernels (glm.cu, irt.cu) and only run in the
// lock-step engine.
#pragma once
#include "warp.cuh"

namespace b200 {

enum ModelKind {
  MODEL_STD_NORMAL = 0,     // x ~ normal(0,1), D free            (ST/test-models/good/mcmc/hmc/common/gauss3D.stan for D=3)
  MODEL_BERNOULLI = 1,      // C/examples/bernoulli/bernoulli.stan (config c1)
  MODEL_EIGHT_SCHOOLS = 2,  // non-centred eight schools           (config c2)
  MODEL_LOGISTIC = 3,       // bernoulli_logit_glm                 (config c3)
  MODEL_IRT2PL = 4,         // hierarchical IRT 2PL                (config c4)
  MODEL_POISSON = 5         // poisson_log_glm                     (config c5)
};

struct ModelData {
  int kind;
  int D;              // unconstrained dimension
  int N;              // observations: bernoulli N / schools J / GLM rows (local shard) / IRT persons
  int K;              // GLM covariates / IRT items
  int ldx;            // GLM: leading dimension of the padded design matrix [1 | X | 0..]
  int sum_y;          // bernoulli: sufficient statistic
  const double* X;    // GLM: N x ldx row-major, column 0 is the intercept column of ones
  const int* yi;      // GLM outcomes (int32)
  const double* yr;   // eight schools y
  const double* sigma;// eight schools sigma
  const signed char* y8;  // IRT responses, persons x items row-major
  const unsigned short* ypk;  // IRT responses packed for the gradient kernels: [persons][8], bit t of word s = y[person][s + 8 t]
};

B200_DEV double dev_log1p_exp(double a) { return a > 0.0 ? a + dev_log1p(dev_exp(-a)) : dev_log1p(dev_exp(a)); }

// M/prim/fun/inv_logit.hpp:51-60
B200_DEV double dev_inv_logit(double u) {
  if (u < 0) {
    double e = dev_exp(u);
    if (u < -36.04365338911715) return e;
    return e / (1 + e);
  }
  return 1 / (1 + dev_exp(-u));
}

// log_sum_exp(a, b), M/prim/fun/log_sum_exp.hpp:51-64 (NaN falls through to the last branch like the reference)
B200_DEV_CALL double dev_log_sum_exp(double a, double b) {
  if (a == -INFINITY) return b;
  if (a == INFINITY && b == INFINITY) return INFINITY;
  if (a > b) return a + dev_log1p_exp(b - a);
  return b + dev_log1p_exp(a - b);
}

// ---- kind 0: standard normal ---------------------------------------------------------------------------------
// normal_lpdf<propto>(x, 0, 1) = -0.5 sum x^2  (M/prim/prob/normal_lpdf.hpp:83-89)
B200_DEV_CALL double lp_grad_std_normal(const ModelData& m, const double* q, double* g, int lane) {
  double s = 0;
  _Pragma("unroll 1") for (int i = lane; i < m.D; i += kWarp) {
    double x = q[i];
    s += x * x;
    g[i] = -x;
  }
  return -0.5 * warp_sum(s);
}

// ---- kind 1: bernoulli (c1) ----------------------------------------------------------------------------------
// theta = inv_logit(u); lub_constrain Jacobian -|u| - 2 log1p_exp(-|u|) (M/rev/constraint/lub_constrain.hpp:100-148);
// bernoulli_lpmf via the sufficient statistic (M/prim/prob/bernoulli_lpmf.hpp:34-114); beta_lpdf(theta|1,1) kernel
// contributes 0*log(theta) + 0*log1m(theta) (M/prim/prob/beta_lpdf.hpp:75-92).  D = 1: every lane computes the
// same scalars, lane 0 owns g[0].
B200_DEV_CALL double lp_grad_bernoulli(const ModelData& m, const double* q, double* g, int lane) {
  double u = q[0];
  int S = m.sum_y, N = m.N;
  double th = dev_inv_logit(u);
  double au = fabs(u);
  double lp = -au - 2.0 * dev_log1p_exp(-au);
  double lth = dev_log(th), l1m = dev_log1p(-th);
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
  if (lane == 0) g[0] = dth * th * (1 - th) + (1 - 2 * th);
  return lp;
}

// ---- kind 2: eight schools, non-centred (c2) -----------------------------------------------------------------
// u = (mu, log tau, theta_tilde[J]).  normal_lpdf.hpp:44-105, cauchy_lpdf.hpp:80-86, lb_constrain (lp += u).
// Element 2+j (school j) is handled by its owner lane.
B200_DEV_CALL double lp_grad_eight_schools(const ModelData& m, const double* q, double* g, int lane) {
  const int J = m.N;
  double mu = q[0], ell = q[1], tau = dev_exp(ell);
  double ss = 0, sq = 0, sr = 0, srt = 0;
  _Pragma("unroll 1") for (int idx = lane; idx < m.D; idx += kWarp) {
    if (idx >= 2) {
      int j = idx - 2;
      double tt = q[idx];
      double theta = mu + tau * tt;
      double sg = m.sigma[j];
      double zs = (m.yr[j] - theta) / sg;
      double r = zs / sg;
      ss += tt * tt;
      sq += zs * zs;
      sr += r;
      srt += r * tt;
      g[idx] = -tt + tau * r;
    }
  }
  (void)J;
  ss = warp_sum(ss);
  sq = warp_sum(sq);
  sr = warp_sum(sr);
  srt = warp_sum(srt);
  double z = mu / 5.0, t5 = tau / 5.0;
  double lp = ell;
  lp += -0.5 * z * z;
  lp += -dev_log1p(t5 * t5);
  lp += -0.5 * ss - 0.5 * sq;
  if (lane == 0) g[0] = -mu / 25.0 + sr;
  if (lane == 1 % kWarp) g[1] = 1.0 - 2.0 * t5 * t5 / (1.0 + t5 * t5) + tau * srt;
  return lp;
}

// dispatch for the inlined (fused-kernel) family.  q and g must be visible to the whole warp (warp_sync before).
B200_DEV double lp_grad_small(const ModelData& m, const double* q, double* g, int lane) {
  switch (m.kind) {
    case MODEL_STD_NORMAL: return lp_grad_std_normal(m, q, g, lane);
    case MODEL_BERNOULLI: return lp_grad_bernoulli(m, q, g, lane);
    default: return lp_grad_eight_schools(m, q, g, lane);
  }
}

// number of output columns of model.write_array (constrained parameters + transformed parameters)
B200_HD int model_num_outputs(int kind, int D, int N) { return kind == MODEL_EIGHT_SCHOOLS ? D + N : D; }

// write_array for one draw: unconstrained q[D] -> out[K]; executed by one thread per draw
B200_HD void model_write_array(int kind, int D, int N, int K, const double* q, double* out) {
  switch (kind) {
    case MODEL_BERNOULLI: {
      double u = q[0], th;
      if (u < 0) {
        double e = exp(u);
        th = (u < -36.04365338911715) ? e : e / (1 + e);
      } else {
        th = 1 / (1 + exp(-u));
      }
      out[0] = th;
      break;
    }
    case MODEL_EIGHT_SCHOOLS: {
      int J = N;
      double mu = q[0], tau = exp(q[1]);
      out[0] = mu;
      out[1] = tau;
      for (int j = 0; j < J; j++) out[2 + j] = q[2 + j];
      for (int j = 0; j < J; j++) out[2 + J + j] = mu + tau * q[2 + j];
      break;
    }
    case MODEL_IRT2PL: {
      int I = N, J = K;
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

// one column of write_array (same arithmetic as model_write_array): lets the on-device posterior summary read a
// constrained output column straight from the stored unconstrained draw without materialising the whole row
B200_HD double model_output_column(int kind, int D, int N, int K, const double* q, int k) {
  switch (kind) {
    case MODEL_BERNOULLI: {
      double out;
      model_write_array(kind, D, N, K, q, &out);
      return out;
    }
    case MODEL_EIGHT_SCHOOLS: {
      const int J = N;
      if (k == 1) return exp(q[1]);
      if (k < 2 + J) return q[k];
      return q[0] + exp(q[1]) * q[k - J];
    }
    case MODEL_IRT2PL: {
      const int I = N, J = K;
      const bool positive = k == 0 || k == 1 + I || (k >= 2 + I && k < 2 + I + J) || k == 3 + I + J;
      return positive ? exp(q[k]) : q[k];
    }
    default:
      return q[k];
  }
}

}  // namespace b200
