// Host-side model object of the B200 engine: the part of the stan::model::model_base contract
// (stan/src/stan/model/model_base.hpp:36-705) that the sampling service and the CSV writer use -- model_name,
// get_param_names / get_dims, constrained_param_names / unconstrained_param_names, num_params_r, transform_inits --
// for the hand-written device model family (include/b200nuts.h enum b200nuts_model; densities: SURVEY.md Appendix E).
// log_prob and write_array run on the device (b200nuts_log_prob_grad, b200nuts_fetch_draws constrained=1): a C++
// virtual cannot be called from a kernel, so "drops into model_base" means same names, dims, transforms and columns.
// The construction hook mirrors the reference's link-time hook
//     stan::model::model_base& new_model(stan::io::var_context&, unsigned int seed, std::ostream*)
// (src/cmdstan/command.hpp:71-73).
#pragma once
#include <cmath>
#include <sstream>
#include <stdexcept>
#include <limits>
#include <string>
#include <vector>

#include <cstdlib>
#include <cstring>

#include "../../include/b200nuts.h"
#include "../../include/b200nuts_model.h"
#include "json.hpp"

namespace b200 {
namespace model {

enum constraint_t { FREE = 0, LOWER0 = 1, UNIT_INTERVAL = 2 };
struct block_t {
  std::string name;
  int64_t size;  // -1: scalar, >= 0: vector of that length
  constraint_t constraint;
  int64_t count() const { return size < 0 ? 1 : size; }
};

class device_model {
 public:
  device_model(const std::string& name, const io::var_context& data) : short_name_(name) {
    auto need = [&](const char* v) {
      if (!data.contains_r(v)) throw std::invalid_argument(std::string("variable does not exist; processing stage=data initialization; variable name=") + v + "; base type=double");
    };
    auto scalar_int = [&](const char* v) {
      need(v);
      return (int64_t)data.vals_r(v)[0];
    };
    if (name == "std_normal") {
      id_ = B200NUTS_STD_NORMAL;
      D_ = data.contains_r("D") ? (int64_t)data.vals_r("D")[0] : 3;
      params_ = {{"x", D_, FREE}};
    } else if (name == "bernoulli") {
      id_ = B200NUTS_BERNOULLI;
      const int64_t N = scalar_int("N");
      need("y");
      yi_ = data.vals_i("y");
      check_size("y", (int64_t)yi_.size(), N);
      for (int v : yi_)
        if (v < 0 || v > 1) throw std::domain_error("bernoulli_model: y is " + std::to_string(v) + ", but must be in the interval [0, 1]");
      params_ = {{"theta", -1, UNIT_INTERVAL}};
    } else if (name == "eight_schools_nc") {
      id_ = B200NUTS_EIGHT_SCHOOLS_NC;
      const int64_t J = scalar_int("J");
      need("y");
      need("sigma");
      yr_ = data.vals_r("y");
      sigma_ = data.vals_r("sigma");
      check_size("y", (int64_t)yr_.size(), J);
      check_size("sigma", (int64_t)sigma_.size(), J);
      for (double s : sigma_)
        if (!(s > 0)) throw std::domain_error("eight_schools_nc_model: sigma must be positive");
      params_ = {{"mu", -1, FREE}, {"tau", -1, LOWER0}, {"theta_tilde", J, FREE}};
      tparams_ = {{"theta", J, FREE}};
    } else if (name == "logistic" || name == "poisson_glm") {
      id_ = name == "logistic" ? B200NUTS_LOGISTIC_GLM : B200NUTS_POISSON_GLM;
      N_ = scalar_int("N");
      K_ = scalar_int("D");
      need("X");
      need("y");
      X_ = data.vals_r("X");
      yi_ = data.vals_i("y");
      check_size("X", (int64_t)X_.size(), N_ * K_);
      check_size("y", (int64_t)yi_.size(), N_);
      params_ = {{"alpha", -1, FREE}, {"beta", K_, FREE}};
    } else if (name == "irt_2pl") {
      id_ = B200NUTS_IRT_2PL;
      N_ = scalar_int("I");
      K_ = scalar_int("J");
      need("y");
      std::vector<int> y = data.vals_i("y");
      check_size("y", (int64_t)y.size(), N_ * K_);
      y8_.assign(y.begin(), y.end());
      params_ = {{"sigma_theta", -1, LOWER0}, {"theta", N_, FREE}, {"sigma_a", -1, LOWER0}, {"a", K_, LOWER0},
                 {"mu_b", -1, FREE},          {"sigma_b", -1, LOWER0}, {"b", K_, FREE}};
    } else if (init_plugin(name, data)) {
      return;
    } else {
      throw std::invalid_argument("unknown model '" + name + "' (built in: std_normal bernoulli eight_schools_nc logistic irt_2pl poisson_glm; "
                                  "plug-in libraries are named in the environment variable B200NUTS_MODEL_LIBRARY)");
    }
    num_params_r_ = 0;
    for (auto& b : params_) num_params_r_ += b.count();
  }
  bool is_plugin() const { return id_ >= B200NUTS_MODEL_PLUGIN_BASE; }

  std::string model_name() const { return short_name_ + "_model"; }
  std::vector<std::string> model_compile_info() const {
    return {"stanc_version = b200-nuts hand-written device model (no stanc)", "stancflags = "};
  }
  int model_id() const { return id_; }
  int64_t num_params_r() const { return num_params_r_; }

  void get_param_names(std::vector<std::string>& names, bool emit_tparams = true, bool = true) const {
    names.clear();
    if (is_plugin()) {  // base names of the output columns ("beta.3" -> "beta"), in order of first appearance
      for (const std::string& n : plugin_outputs_) {
        const std::string b = n.substr(0, n.find('.'));
        if (names.empty() || names.back() != b) names.push_back(b);
      }
      return;
    }
    for (auto& b : params_) names.push_back(b.name);
    if (emit_tparams)
      for (auto& b : tparams_) names.push_back(b.name);
  }
  void get_dims(std::vector<std::vector<size_t>>& dimss, bool emit_tparams = true, bool = true) const {
    dimss.clear();
    if (is_plugin()) {  // run lengths of equal base names (a plug-in reports flat output names only)
      std::string last;
      for (const std::string& n : plugin_outputs_) {
        const std::string b = n.substr(0, n.find('.'));
        const bool indexed = n.find('.') != std::string::npos;
        if (dimss.empty() || b != last) dimss.push_back(indexed ? std::vector<size_t>{1} : std::vector<size_t>{});
        else dimss.back()[0]++;
        last = b;
      }
      return;
    }
    auto push = [&](const block_t& b) { dimss.push_back(b.size < 0 ? std::vector<size_t>{} : std::vector<size_t>{(size_t)b.size}); };
    for (auto& b : params_) push(b);
    if (emit_tparams)
      for (auto& b : tparams_) push(b);
  }
  // names are appended (as model_base::constrained_param_names does): "theta", "beta.1", ...
  void constrained_param_names(std::vector<std::string>& names, bool emit_tparams = true, bool = true) const {
    if (is_plugin()) {
      names.insert(names.end(), plugin_outputs_.begin(), plugin_outputs_.end());
      return;
    }
    for (auto& b : params_) expand(b, names);
    if (emit_tparams)
      for (auto& b : tparams_) expand(b, names);
  }
  void unconstrained_param_names(std::vector<std::string>& names, bool = true, bool = true) const {
    if (is_plugin()) {
      for (int64_t i = 0; i < num_params_r_; i++) names.push_back("u." + std::to_string(i + 1));
      return;
    }
    for (auto& b : params_) expand(b, names);
  }

  // constrained initial values -> unconstrained vector (model_base::transform_inits; lb_free = log,
  // lub_free(0,1) = logit: stan/math/prim/constraint/lb_free.hpp, lub_free.hpp).  Parameters that are absent are left to the engine.
  void transform_inits(const io::var_context& init, std::vector<double>& params_r) const {
    params_r.clear();
    if (is_plugin()) {  // the model's own transform_inits (include/b200nuts_model.h), through the C ABI
      plugin_transform_inits(init, params_r);
      return;
    }
    for (auto& b : params_) {
      if (!init.contains_r(b.name)) {
        // partial inits (initialize.hpp:93-101: chained_var_context(init, random_context)): the engine draws this parameter like
        // the reference's random context does; a quiet NaN whose payload carries the transform marks "not given" (csrc/nuts.cuh)
        const unsigned long long bits = 0x7ff8000000000000ull | (unsigned long long)(b.constraint == LOWER0 ? 1 : (b.constraint == UNIT_INTERVAL ? 2 : 0));
        double nan_code;
        std::memcpy(&nan_code, &bits, sizeof nan_code);
        for (int64_t i = 0; i < b.count(); i++) params_r.push_back(nan_code);
        continue;
      }
      const std::vector<double>& v = init.vals_r(b.name);
      if ((int64_t)v.size() != b.count()) throw std::invalid_argument("mismatch in number dimensions declared and found in context; processing stage=parameter initialization; variable name=" + b.name);
      for (double x : v) {
        if (b.constraint == LOWER0) {
          if (!(x >= 0)) throw std::domain_error("lb_free: Bounded variable is " + std::to_string(x) + ", but must be greater than or equal to 0");
          params_r.push_back(std::log(x));
        } else if (b.constraint == UNIT_INTERVAL) {
          if (!(x >= 0 && x <= 1)) throw std::domain_error("lub_free: Bounded variable is " + std::to_string(x) + ", but must be in the interval [0, 1]");
          params_r.push_back(std::log(x / (1 - x)));
        } else {
          params_r.push_back(x);
        }
      }
    }
  }

  // hands the data to an engine through the C ABI
  int upload(b200nuts_engine* e) const {
    int rc = 0;
    int64_t d1[1], d2[2];
    if (is_plugin()) return upload_all(e, plugin_data_);
    switch (id_) {
      case B200NUTS_STD_NORMAL: {
        int32_t D = (int32_t)D_;
        d1[0] = 1;
        return b200nuts_set_data_i32(e, "D", &D, d1, 1);
      }
      case B200NUTS_BERNOULLI:
        d1[0] = (int64_t)yi_.size();
        return b200nuts_set_data_i32(e, "y", yi_.data(), d1, 1);
      case B200NUTS_EIGHT_SCHOOLS_NC:
        d1[0] = (int64_t)yr_.size();
        if ((rc = b200nuts_set_data_f64(e, "y", yr_.data(), d1, 1))) return rc;
        return b200nuts_set_data_f64(e, "sigma", sigma_.data(), d1, 1);
      case B200NUTS_LOGISTIC_GLM:
      case B200NUTS_POISSON_GLM:
        d2[0] = N_;
        d2[1] = K_;
        d1[0] = N_;
        if ((rc = b200nuts_set_data_f64(e, "X", X_.data(), d2, 2))) return rc;
        return b200nuts_set_data_i32(e, "y", yi_.data(), d1, 1);
      case B200NUTS_IRT_2PL:
        d2[0] = N_;
        d2[1] = K_;
        return b200nuts_set_data_i8(e, "y", y8_.data(), d2, 2);
    }
    return B200NUTS_USAGE;
  }

 private:
  // every variable of the data context, as given: integer-valued ones as i32, the rest as f64 (the model's create() validates)
  static int upload_all(b200nuts_engine* e, const io::var_context& data) {
    for (const std::string& n : data.names()) {
      const std::vector<int64_t>& dims = data.dims_r(n);
      const int64_t one = 1;
      const int64_t* dp = dims.empty() ? &one : dims.data();
      const int nd = dims.empty() ? 1 : (int)dims.size();
      int rc;
      if (data.contains_i(n)) {
        std::vector<int> v = data.vals_i(n);
        rc = b200nuts_set_data_i32(e, n.c_str(), v.data(), dp, nd);
      } else {
        rc = b200nuts_set_data_f64(e, n.c_str(), data.vals_r(n).data(), dp, nd);
      }
      if (rc) return rc;
    }
    return 0;
  }
  // A model that is not built in: look it up among the plug-in libraries of B200NUTS_MODEL_LIBRARY (':'-separated paths,
  // include/b200nuts_model.h), construct it once on device 0 to learn its dimensions and output names.
  bool init_plugin(const std::string& name, const io::var_context& data) {
    if (const char* libs = std::getenv("B200NUTS_MODEL_LIBRARY")) {
      std::stringstream ss(libs);
      std::string path;
      while (std::getline(ss, path, ':'))
        if (!path.empty() && b200nuts_load_model_library(path.c_str()) < 0) throw std::invalid_argument(b200nuts_last_error(nullptr));
    }
    const int id = b200nuts_find_model(name.c_str());
    if (id < B200NUTS_MODEL_PLUGIN_BASE) return false;
    id_ = id;
    plugin_data_ = data;
    b200nuts_config cfg;
    b200nuts_default_config(&cfg);
    cfg.model = id;
    b200nuts_engine* e = nullptr;
    if (b200nuts_create(&cfg, &e)) throw std::runtime_error(b200nuts_last_error(nullptr));
    int64_t D = 0, K = 0;
    int rc = upload_all(e, data);
    if (!rc) rc = b200nuts_num_params(e, &D, &K);
    if (rc) {
      const std::string msg = b200nuts_last_error(e);
      b200nuts_destroy(e);
      if (rc == B200NUTS_DATAERR) throw std::domain_error(msg);
      throw std::invalid_argument(msg);
    }
    char buf[256];
    for (int64_t k = 0; k < K; k++) {
      b200nuts_output_name(e, k, buf, (int)sizeof buf);
      plugin_outputs_.push_back(buf);
    }
    b200nuts_destroy(e);
    num_params_r_ = D;
    return true;
  }
  // constrained values of the init context, laid out like the plug-in's output columns (NaN = not given), then
  // b200nuts_transform_inits.  The context's variables are matched to the columns by base name ("beta" -> "beta.1", ...).
  void plugin_transform_inits(const io::var_context& init, std::vector<double>& params_r) const {
    const size_t K = plugin_outputs_.size();
    std::vector<double> x(K, std::numeric_limits<double>::quiet_NaN());
    size_t k = 0;
    while (k < K) {
      const std::string b = plugin_outputs_[k].substr(0, plugin_outputs_[k].find('.'));
      size_t cnt = 0;
      while (k + cnt < K && plugin_outputs_[k + cnt].substr(0, plugin_outputs_[k + cnt].find('.')) == b) cnt++;
      if (init.contains_r(b)) {
        const std::vector<double>& v = init.vals_r(b);
        if (v.size() != cnt) throw std::invalid_argument("mismatch in number dimensions declared and found in context; processing stage=parameter initialization; variable name=" + b);
        std::copy(v.begin(), v.end(), x.begin() + k);
      }
      k += cnt;
    }
    if (!plugin_engine_) {
      b200nuts_config cfg;
      b200nuts_default_config(&cfg);
      cfg.model = id_;
      if (b200nuts_create(&cfg, &plugin_engine_)) throw std::runtime_error(b200nuts_last_error(nullptr));
      if (upload_all(plugin_engine_, plugin_data_)) throw std::invalid_argument(b200nuts_last_error(plugin_engine_));
    }
    params_r.assign((size_t)num_params_r_, 0.0);
    const int rc = b200nuts_transform_inits(plugin_engine_, x.data(), 1, params_r.data());
    if (rc == B200NUTS_CONFIG) throw std::domain_error(b200nuts_last_error(plugin_engine_));
    if (rc) throw std::invalid_argument(b200nuts_last_error(plugin_engine_));
  }
  static void expand(const block_t& b, std::vector<std::string>& names) {
    if (b.size < 0) names.push_back(b.name);
    else
      for (int64_t i = 1; i <= b.size; i++) names.push_back(b.name + "." + std::to_string(i));
  }
  void check_size(const char* v, int64_t got, int64_t want) const {
    if (got != want)
      throw std::invalid_argument(std::string("mismatch in dimension declared and found in context; processing stage=data initialization; variable name=") + v +
                                  "; dims declared=(" + std::to_string(want) + "); dims found=(" + std::to_string(got) + ")");
  }
  std::string short_name_;
  int id_ = 0;
  int64_t D_ = 0, N_ = 0, K_ = 0, num_params_r_ = 0;
  std::vector<block_t> params_, tparams_;
  std::vector<double> X_, yr_, sigma_;
  std::vector<int> yi_;
  std::vector<int8_t> y8_;
  io::var_context plugin_data_;
  std::vector<std::string> plugin_outputs_;
  mutable b200nuts_engine* plugin_engine_ = nullptr;  // holds the plug-in's model object for transform_inits (created on first use)

 public:
  ~device_model() {
    if (plugin_engine_) b200nuts_destroy(plugin_engine_);
  }
  device_model(const device_model&) = delete;
  device_model& operator=(const device_model&) = delete;
};

// the reference's link-time hook, with the model selected by name instead of by linking a stanc-generated class
inline device_model& new_model(const std::string& name, const io::var_context& data, unsigned int /*seed*/, std::ostream* /*msg*/) {
  return *new device_model(name, data);
}

}  // namespace model
}  // namespace b200
