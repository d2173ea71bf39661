// TEST INFRASTRUCTURE ONLY (CPU oracle).  Hierarchical IRT 2PL (config c4; SURVEY.md Appendix E, posteriordb irt_2pl):
//   parameters { real<lower=0> sigma_theta; vector[I] theta; real<lower=0> sigma_a; vector<lower=0>[J] a;
//                real mu_b; real<lower=0> sigma_b; vector[J] b; }
//   model { sigma_theta ~ cauchy(0,2); theta ~ normal(0,sigma_theta); sigma_a ~ cauchy(0,2); a ~ lognormal(0,sigma_a);
//           mu_b ~ normal(0,5); sigma_b ~ cauchy(0,2); b ~ normal(mu_b,sigma_b);
//           for (j in 1:J) y[,j] ~ bernoulli_logit(a[j] * (theta - b[j])); }
// Data: I (persons), J (items), y as an I x J 0/1 matrix -- var_context "y" (column-major ints) or the raw file
// $B200_ORACLE_DATA_BIN = y[I][J] int8 row-major.
#include "model_common.h"
class irt_2pl_model final : public b200_oracle::oracle_model<irt_2pl_model> {
  int I, J;
  std::vector<std::vector<int>> ycol;  // ycol[j][i]
 public:
  static const char* name() { return "irt_2pl_model"; }
  irt_2pl_model(stan::io::var_context& ctx, unsigned int, std::ostream*) {
    I = ctx.vals_i("I")[0];
    J = ctx.vals_i("J")[0];
    ycol.assign(J, std::vector<int>(I));
    if (ctx.contains_i("y")) {
      std::vector<int> yv = ctx.vals_i("y");  // column-major I x J
      stan::math::check_size_match("irt_2pl", "y", yv.size(), "I*J", static_cast<size_t>(I) * J);
      for (int j = 0; j < J; ++j)
        for (int i = 0; i < I; ++i) ycol[j][i] = yv[static_cast<size_t>(j) * I + i];
    } else {
      std::vector<char> buf = b200_oracle::read_data_bin();
      stan::math::check_size_match("irt_2pl", "data_bin", buf.size(), "I*J", static_cast<size_t>(I) * J);
      for (int i = 0; i < I; ++i)
        for (int j = 0; j < J; ++j) ycol[j][i] = buf[static_cast<size_t>(i) * J + j];
    }
    for (int j = 0; j < J; ++j) stan::math::check_bounded("irt_2pl", "y", ycol[j], 0, 1);
    params_ = {{"sigma_theta", -1, b200_oracle::LOWER0}, {"theta", I, b200_oracle::FREE},
               {"sigma_a", -1, b200_oracle::LOWER0},     {"a", J, b200_oracle::LOWER0},
               {"mu_b", -1, b200_oracle::FREE},          {"sigma_b", -1, b200_oracle::LOWER0},
               {"b", J, b200_oracle::FREE}};
    finish_blocks();
  }
  template <bool propto__, bool jacobian__, typename T__>
  T__ lp(const Eigen::Matrix<T__, -1, 1>& params_r__, std::ostream* = nullptr) const {
    using vec_t = Eigen::Matrix<T__, -1, 1>;
    T__ lp__(0.0);
    stan::math::accumulator<T__> lp_accum__;
    stan::io::deserializer<T__> in__(params_r__, std::vector<int>{});
    T__ sigma_theta = in__.template read_constrain_lb<T__, jacobian__>(0, lp__);
    vec_t theta = in__.template read<vec_t>(I);
    T__ sigma_a = in__.template read_constrain_lb<T__, jacobian__>(0, lp__);
    vec_t a = in__.template read_constrain_lb<vec_t, jacobian__>(0, lp__, J);
    T__ mu_b = in__.template read<T__>();
    T__ sigma_b = in__.template read_constrain_lb<T__, jacobian__>(0, lp__);
    vec_t b = in__.template read<vec_t>(J);
    lp_accum__.add(stan::math::cauchy_lpdf<propto__>(sigma_theta, 0, 2));
    lp_accum__.add(stan::math::normal_lpdf<propto__>(theta, 0, sigma_theta));
    lp_accum__.add(stan::math::cauchy_lpdf<propto__>(sigma_a, 0, 2));
    lp_accum__.add(stan::math::lognormal_lpdf<propto__>(a, 0, sigma_a));
    lp_accum__.add(stan::math::normal_lpdf<propto__>(mu_b, 0, 5));
    lp_accum__.add(stan::math::cauchy_lpdf<propto__>(sigma_b, 0, 2));
    lp_accum__.add(stan::math::normal_lpdf<propto__>(b, mu_b, sigma_b));
    for (int j = 0; j < J; ++j)
      lp_accum__.add(stan::math::bernoulli_logit_lpmf<propto__>(
          ycol[j], stan::math::multiply(a[j], stan::math::subtract(theta, b[j]))));
    lp_accum__.add(lp__);
    return lp_accum__.sum();
  }
  void tparams(const std::vector<double>&, std::vector<double>&) const {}
};
B200_ORACLE_DEFINE_HOOKS(irt_2pl_model)
