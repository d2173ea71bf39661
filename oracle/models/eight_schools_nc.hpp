// TEST INFRASTRUCTURE ONLY (CPU oracle).  Non-centred eight schools (config c2):
//   parameters { real mu; real<lower=0> tau; vector[J] theta_tilde; }
//   transformed parameters { vector[J] theta = mu + tau * theta_tilde; }
//   model { mu ~ normal(0,5); tau ~ cauchy(0,5); theta_tilde ~ std_normal(); y ~ normal(theta, sigma); }
#include "model_common.h"
class eight_schools_nc_model final : public b200_oracle::oracle_model<eight_schools_nc_model> {
  int J;
  Eigen::VectorXd y, sigma;
 public:
  static const char* name() { return "eight_schools_nc_model"; }
  eight_schools_nc_model(stan::io::var_context& ctx, unsigned int, std::ostream*) {
    J = ctx.vals_i("J")[0];
    std::vector<double> yv = ctx.vals_r("y"), sv = ctx.vals_r("sigma");
    stan::math::check_size_match("eight_schools", "y", yv.size(), "J", J);
    stan::math::check_size_match("eight_schools", "sigma", sv.size(), "J", J);
    y = Eigen::Map<Eigen::VectorXd>(yv.data(), J);
    sigma = Eigen::Map<Eigen::VectorXd>(sv.data(), J);
    stan::math::check_positive("eight_schools", "sigma", sigma);
    params_ = {{"mu", -1, b200_oracle::FREE}, {"tau", -1, b200_oracle::LOWER0}, {"theta_tilde", J, b200_oracle::FREE}};
    tparams_ = {{"theta", J, b200_oracle::FREE}};
    finish_blocks();
  }
  template <bool propto__, bool jacobian__, typename T__>
  T__ lp(const Eigen::Matrix<T__, -1, 1>& params_r__, std::ostream* = nullptr) const {
    using vec_t = Eigen::Matrix<T__, -1, 1>;
    T__ lp__(0.0);
    stan::math::accumulator<T__> lp_accum__;
    stan::io::deserializer<T__> in__(params_r__, std::vector<int>{});
    T__ mu = in__.template read<T__>();
    T__ tau = in__.template read_constrain_lb<T__, jacobian__>(0, lp__);
    vec_t theta_tilde = in__.template read<vec_t>(J);
    vec_t theta = stan::math::add(mu, stan::math::multiply(tau, theta_tilde));
    lp_accum__.add(stan::math::normal_lpdf<propto__>(mu, 0, 5));
    lp_accum__.add(stan::math::cauchy_lpdf<propto__>(tau, 0, 5));
    lp_accum__.add(stan::math::std_normal_lpdf<propto__>(theta_tilde));
    lp_accum__.add(stan::math::normal_lpdf<propto__>(y, theta, sigma));
    lp_accum__.add(lp__);
    return lp_accum__.sum();
  }
  void tparams(const std::vector<double>& p, std::vector<double>& out) const {
    for (int j = 0; j < J; ++j) out.push_back(p[0] + p[1] * p[2 + j]);
  }
};
B200_ORACLE_DEFINE_HOOKS(eight_schools_nc_model)
