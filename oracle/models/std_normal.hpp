// TEST INFRASTRUCTURE ONLY (CPU oracle).  D-dimensional standard normal; D = 3 reproduces
// ST/test-models/good/mcmc/hmc/common/gauss3D.stan (the model of unit_e_nuts_test.cpp's golden values).
#include "model_common.h"
class std_normal_model final : public b200_oracle::oracle_model<std_normal_model> {
  int D;
 public:
  static const char* name() { return "std_normal_model"; }
  std_normal_model(stan::io::var_context& ctx, unsigned int, std::ostream*) {
    D = ctx.contains_i("D") ? ctx.vals_i("D")[0] : 3;
    params_ = {{"x", D, b200_oracle::FREE}};
    finish_blocks();
  }
  template <bool propto__, bool jacobian__, typename T__>
  T__ lp(const Eigen::Matrix<T__, -1, 1>& params_r__, std::ostream* = nullptr) const {
    stan::math::accumulator<T__> lp_accum__;
    stan::io::deserializer<T__> in__(params_r__, std::vector<int>{});
    std::vector<T__> x = in__.template read<std::vector<T__>>(D);
    lp_accum__.add(stan::math::normal_lpdf<propto__>(x, 0, 1));
    return lp_accum__.sum();
  }
  void tparams(const std::vector<double>&, std::vector<double>&) const {}
};
B200_ORACLE_DEFINE_HOOKS(std_normal_model)
