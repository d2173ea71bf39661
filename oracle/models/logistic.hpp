// TEST INFRASTRUCTURE ONLY (CPU oracle).  Logistic regression GLM (config c3):
//   parameters { real alpha; vector[D] beta; }
//   model { alpha ~ normal(0,2.5); beta ~ normal(0,1); y ~ bernoulli_logit_glm(X, alpha, beta); }
// Data: var_context N, D and (X row-major as matrix[N,D], y) -- or, when X is absent, the raw file
// $B200_ORACLE_DATA_BIN = X[N][D] float64 row-major followed by y[N] int32.
#include "model_common.h"
class logistic_model final : public b200_oracle::oracle_model<logistic_model> {
  int N, D;
  Eigen::MatrixXd X;
  std::vector<int> y;
 public:
  static const char* name() { return "logistic_model"; }
  logistic_model(stan::io::var_context& ctx, unsigned int, std::ostream*) {
    N = ctx.vals_i("N")[0];
    D = ctx.vals_i("D")[0];
    X.resize(N, D);
    if (ctx.contains_r("X")) {
      std::vector<double> xv = ctx.vals_r("X");  // column-major (Stan convention for matrix data)
      stan::math::check_size_match("logistic", "X", xv.size(), "N*D", static_cast<size_t>(N) * D);
      X = Eigen::Map<Eigen::MatrixXd>(xv.data(), N, D);
      y = ctx.vals_i("y");
    } else {
      std::vector<char> buf = b200_oracle::read_data_bin();
      stan::math::check_size_match("logistic", "data_bin", buf.size(), "expected", static_cast<size_t>(N) * D * 8 + N * 4);
      const double* xp = reinterpret_cast<const double*>(buf.data());
      for (int i = 0; i < N; ++i)
        for (int j = 0; j < D; ++j) X(i, j) = xp[static_cast<size_t>(i) * D + j];
      const int* yp = reinterpret_cast<const int*>(buf.data() + static_cast<size_t>(N) * D * 8);
      y.assign(yp, yp + N);
    }
    stan::math::check_size_match("logistic", "y", y.size(), "N", N);
    stan::math::check_bounded("logistic", "y", y, 0, 1);
    params_ = {{"alpha", -1, b200_oracle::FREE}, {"beta", D, b200_oracle::FREE}};
    finish_blocks();
  }
  template <bool propto__, bool jacobian__, typename T__>
  T__ lp(const Eigen::Matrix<T__, -1, 1>& params_r__, std::ostream* = nullptr) const {
    using vec_t = Eigen::Matrix<T__, -1, 1>;
    stan::math::accumulator<T__> lp_accum__;
    stan::io::deserializer<T__> in__(params_r__, std::vector<int>{});
    T__ alpha = in__.template read<T__>();
    vec_t beta = in__.template read<vec_t>(D);
    lp_accum__.add(stan::math::normal_lpdf<propto__>(alpha, 0, 2.5));
    lp_accum__.add(stan::math::normal_lpdf<propto__>(beta, 0, 1));
    lp_accum__.add(stan::math::bernoulli_logit_glm_lpmf<propto__>(y, X, alpha, beta));
    return lp_accum__.sum();
  }
  void tparams(const std::vector<double>&, std::vector<double>&) const {}
};
B200_ORACLE_DEFINE_HOOKS(logistic_model)
