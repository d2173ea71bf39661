// TEST INFRASTRUCTURE ONLY -- part of the CPU oracle, never linked into the product.
//
// stanc (the Stan -> C++ transpiler) is not available offline, so every oracle model is a hand-written
// stan::model::model_base_crtp<M> class that calls the *reference's own* stan::math densities and
// stan::io::deserializer transforms -- the arithmetic (and the reverse-mode AD gradient) is 100 % the
// reference's.  This header factors the boilerplate stanc would emit (names, dims, transform_inits,
// write_array, unconstrain_array, new_model hook) out of the per-model files.
//
// Contract followed: S/model/model_base.hpp:36-705, S/model/model_base_crtp.hpp,
// link-time hooks C/src/cmdstan/command.hpp:71-73.
#ifndef B200_ORACLE_MODEL_COMMON_H
#define B200_ORACLE_MODEL_COMMON_H

#include <stan/model/model_header.hpp>

#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

namespace b200_oracle {

enum constraint_t { FREE = 0, LOWER0 = 1, UNIT_INTERVAL = 2 };

struct block_t {
  std::string name;
  int size;  // -1: scalar, >=0: vector of that length
  constraint_t constraint;
  int count() const { return size < 0 ? 1 : size; }
};

// M must provide:
//   static const char* name();
//   template <bool propto, bool jacobian, typename T> T lp(const Eigen::Matrix<T,-1,1>& theta, std::ostream*) const;
//   void tparams(const std::vector<double>& constrained_params, std::vector<double>& out) const;  (may append nothing)
// and fill params_ / tparams_ in its constructor, then call finish_blocks().
template <class M>
class oracle_model : public stan::model::model_base_crtp<M> {
 public:
  oracle_model() : stan::model::model_base_crtp<M>(0) {}

  std::string model_name() const final { return M::name(); }
  std::vector<std::string> model_compile_info() const noexcept {
    return std::vector<std::string>{"stanc_version = hand-written oracle model (no stanc)", "stancflags = "};
  }

  void get_param_names(std::vector<std::string>& names, bool emit_tparams = true, bool emit_gqs = true) const {
    names.clear();
    for (const auto& b : params_) names.push_back(b.name);
    if (emit_tparams)
      for (const auto& b : tparams_) names.push_back(b.name);
  }
  void get_dims(std::vector<std::vector<size_t>>& dimss, bool emit_tparams = true, bool emit_gqs = true) const {
    dimss.clear();
    auto push = [&](const block_t& b) {
      if (b.size < 0) dimss.push_back(std::vector<size_t>{});
      else dimss.push_back(std::vector<size_t>{static_cast<size_t>(b.size)});
    };
    for (const auto& b : params_) push(b);
    if (emit_tparams)
      for (const auto& b : tparams_) push(b);
  }
  void constrained_param_names(std::vector<std::string>& names, bool emit_tparams = true, bool emit_gqs = true) const {
    auto push = [&](const block_t& b) {
      if (b.size < 0) names.push_back(b.name);
      else for (int i = 1; i <= b.size; ++i) names.push_back(b.name + "." + std::to_string(i));
    };
    for (const auto& b : params_) push(b);
    if (emit_tparams)
      for (const auto& b : tparams_) push(b);
  }
  void unconstrained_param_names(std::vector<std::string>& names, bool emit_tparams = true, bool emit_gqs = true) const {
    constrained_param_names(names, emit_tparams, emit_gqs);
  }

  // ---- log_prob (Eigen and std::vector forms) ----
  template <bool propto, bool jacobian, typename T>
  T log_prob(Eigen::Matrix<T, -1, 1>& params_r, std::ostream* msgs = nullptr) const {
    return static_cast<const M*>(this)->template lp<propto, jacobian, T>(params_r, msgs);
  }
  template <bool propto, bool jacobian, typename T>
  T log_prob(std::vector<T>& params_r, std::vector<int>& params_i, std::ostream* msgs = nullptr) const {
    Eigen::Matrix<T, -1, 1> v = Eigen::Map<Eigen::Matrix<T, -1, 1>>(params_r.data(), params_r.size());
    return static_cast<const M*>(this)->template lp<propto, jacobian, T>(v, msgs);
  }

  // ---- constrained output ----
  void constrain(const double* u, std::vector<double>& out, bool emit_tparams) const {
    out.clear();
    size_t pos = 0;
    for (const auto& b : params_)
      for (int i = 0; i < b.count(); ++i, ++pos) {
        double x = u[pos];
        if (b.constraint == LOWER0) x = stan::math::lb_constrain(x, 0);
        else if (b.constraint == UNIT_INTERVAL) x = stan::math::lub_constrain(x, 0, 1);
        out.push_back(x);
      }
    if (emit_tparams) {
      std::vector<double> cp(out);
      static_cast<const M*>(this)->tparams(cp, out);
    }
  }
  template <typename RNG>
  void write_array(RNG& base_rng, Eigen::VectorXd& params_r, Eigen::VectorXd& vars, bool emit_tparams = true,
                   bool emit_gqs = true, std::ostream* msgs = nullptr) const {
    std::vector<double> out;
    constrain(params_r.data(), out, emit_tparams);
    vars = Eigen::Map<Eigen::VectorXd>(out.data(), out.size());
  }
  template <typename RNG>
  void write_array(RNG& base_rng, std::vector<double>& params_r, std::vector<int>& params_i, std::vector<double>& vars,
                   bool emit_tparams = true, bool emit_gqs = true, std::ostream* msgs = nullptr) const {
    constrain(params_r.data(), vars, emit_tparams);
  }

  // ---- inits ----
  void free_block(const block_t& b, const std::vector<double>& vals, std::vector<double>& out) const {
    for (int i = 0; i < b.count(); ++i) {
      double x = vals.at(i);
      if (b.constraint == LOWER0) x = stan::math::lb_free(x, 0);
      else if (b.constraint == UNIT_INTERVAL) x = stan::math::lub_free(x, 0, 1);
      out.push_back(x);
    }
  }
  void transform_inits(const stan::io::var_context& context, std::vector<int>& params_i, std::vector<double>& params_r,
                       std::ostream* msgs = nullptr) const final {
    params_r.clear();
    for (const auto& b : params_) {
      if (!context.contains_r(b.name)) throw std::runtime_error("variable " + b.name + " missing");
      free_block(b, context.vals_r(b.name), params_r);
    }
  }
  void transform_inits(const stan::io::var_context& context, Eigen::VectorXd& params_r,
                       std::ostream* msgs = nullptr) const final {
    std::vector<int> pi;
    std::vector<double> pr;
    transform_inits(context, pi, pr, msgs);
    params_r = Eigen::Map<Eigen::VectorXd>(pr.data(), pr.size());
  }
  void unconstrain_array(const std::vector<double>& constrained, std::vector<double>& params_r,
                         std::ostream* msgs = nullptr) const final {
    params_r.clear();
    size_t pos = 0;
    for (const auto& b : params_) {
      std::vector<double> vals(constrained.begin() + pos, constrained.begin() + pos + b.count());
      free_block(b, vals, params_r);
      pos += b.count();
    }
  }
  void unconstrain_array(const Eigen::VectorXd& constrained, Eigen::VectorXd& params_r,
                         std::ostream* msgs = nullptr) const final {
    std::vector<double> c(constrained.data(), constrained.data() + constrained.size()), pr;
    unconstrain_array(c, pr, msgs);
    params_r = Eigen::Map<Eigen::VectorXd>(pr.data(), pr.size());
  }

 protected:
  void finish_blocks() {
    size_t n = 0;
    for (const auto& b : params_) n += b.count();
    this->num_params_r__ = n;
  }
  std::vector<block_t> params_, tparams_;
};

// Large synthetic data sets are handed over as a raw little-endian file (the JSON would be ~200 MB of text):
// env B200_ORACLE_DATA_BIN = path; layout is model-specific (see logistic.hpp / poisson_glm.hpp / irt_2pl.hpp).
inline std::vector<char> read_data_bin() {
  const char* p = std::getenv("B200_ORACLE_DATA_BIN");
  if (!p) throw std::runtime_error("data not in var_context and B200_ORACLE_DATA_BIN is not set");
  FILE* f = std::fopen(p, "rb");
  if (!f) throw std::runtime_error(std::string("cannot open ") + p);
  std::fseek(f, 0, SEEK_END);
  long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  std::vector<char> buf(n);
  if (std::fread(buf.data(), 1, n, f) != static_cast<size_t>(n)) throw std::runtime_error("short read");
  std::fclose(f);
  return buf;
}

}  // namespace b200_oracle

#define B200_ORACLE_DEFINE_HOOKS(MODEL)                                                                          \
  stan::model::model_base& new_model(stan::io::var_context& data_context, unsigned int seed, std::ostream* msg) { \
    MODEL* m = new MODEL(data_context, seed, msg);                                                               \
    return *m;                                                                                                   \
  }                                                                                                              \
  stan::math::profile_map& get_stan_profile_data() {                                                             \
    static stan::math::profile_map m;                                                                            \
    return m;                                                                                                    \
  }

#endif
