// TEST INFRASTRUCTURE ONLY (CPU oracle).  Hand-written equivalent of C/examples/bernoulli/bernoulli.stan (config c1).
#include "model_common.h"
class bernoulli_model final : public b200_oracle::oracle_model<bernoulli_model> {
  int N;
  std::vector<int> y;
 public:
  static const char* name() { return "bernoulli_model"; }
  bernoulli_model(stan::io::var_context& ctx, unsigned int, std::ostream*) {
    N = ctx.vals_i("N")[0];
    y = ctx.vals_i("y");
    stan::math::check_size_match("bernoulli", "y", y.size(), "N", N);
    stan::math::check_bounded("bernoulli", "y", y, 0, 1);
    params_ = {{"theta", -1, b200_oracle::UNIT_INTERVAL}};
    finish_blocks();
  }
  template <bool propto__, bool jacobian__, typename T__>
  T__ lp(const Eigen::Matrix<T__, -1, 1>& params_r__, std::ostream* = nullptr) const {
    T__ lp__(0.0);
    stan::math::accumulator<T__> lp_accum__;
    stan::io::deserializer<T__> in__(params_r__, std::vector<int>{});
    T__ theta = in__.template read_constrain_lub<T__, jacobian__>(0, 1, lp__);
    lp_accum__.add(stan::math::beta_lpdf<propto__>(theta, 1, 1));
    lp_accum__.add(stan::math::bernoulli_lpmf<propto__>(y, theta));
    lp_accum__.add(lp__);
    return lp_accum__.sum();
  }
  void tparams(const std::vector<double>&, std::vector<double>&) const {}
};
B200_ORACLE_DEFINE_HOOKS(bernoulli_model)
