// b200::services::sample::hmc_nuts_diag_e_adapt -- host shim with the parameter list of the reference service
//
//   stan::services::sample::hmc_nuts_diag_e_adapt(model, num_chains, init, init_inv_metric, random_seed, init_chain_id,
//       init_radius, num_warmup, num_samples, num_thin, save_warmup, refresh, stepsize, stepsize_jitter, max_depth,
//       delta, gamma, kappa, t0, init_buffer, term_buffer, window, interrupt, logger, init_writer, sample_writer,
//       diagnostic_writer, metric_writer)           stan/src/stan/services/sample/hmc_nuts_diag_e_adapt.hpp:328-407
//
// so that the call site in cmdstan::command() (src/cmdstan/command.hpp:673-688) changes by one namespace
// (INTEGRATION.md).  Everything the reference does per chain inside that call -- create_rng, initialize, sampler
// set-up, run_adaptive_sampler (init_stepsize, warm-up, adaptation block, sampling, timing), mcmc_writer rows -- happens
// here for ALL chains at once: the numerical work on the GPU(s) through the C ABI (include/b200nuts.h), the text
// output on the host in the reference's order and format:
//
//   header row            mcmc_writer::write_sample_names       services/util/mcmc_writer.hpp:64-78
//   draws                 mcmc_writer::write_sample_params      :98-135   (stream precision = sig_figs)
//   "Adaptation terminated" / "Step size = " / metric          :144-146, mcmc/hmc/base_hmc.hpp:46-65,
//                                                                 mcmc/hmc/hamiltonians/diag_e_point.hpp:44-52
//   timing footer         mcmc_writer::write_timing             :198-218
//   progress messages     util::generate_transitions            services/util/generate_transitions.hpp:52-69
//   metric JSON           base_hmc::write_sampler_state_struct  mcmc/hmc/base_hmc.hpp:67-73
//
// Differences, by construction: chains advance together, so rows are written after each phase rather than streamed;
// the interrupt callback runs once per refresh interval; elapsed times are those of the batched run (identical in every
// chain's footer); the diagnostic writer receives the same header but no per-draw momenta (they never leave the GPU).
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <iomanip>
#include <memory>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b200nuts.h"
#include "callbacks.hpp"
#include "json.hpp"
#include "model.hpp"

namespace b200 {
namespace services {

namespace error_codes {  // stan/src/stan/services/error_codes.hpp
enum { OK = 0, USAGE = 64, DATAERR = 65, NOINPUT = 66, SOFTWARE = 70, CONFIG = 78 };
}

// GPU placement of the chains (no counterpart in the reference: it runs chains as TBB tasks)
struct device_options {
  std::vector<int> devices = {0};  // chains are split contiguously over these CUDA devices
  int engine = B200NUTS_ENGINE_AUTO;
  int format_threads = 8;          // host threads used to format CSV rows (per-chain files are independent)
  bool unit_e = false;             // identity metric (the hmc_nuts_unit_e* services below set it)
  bool dense_e = false;            // dense metric (the hmc_nuts_dense_e* services below set it)
  bool static_hmc = false;         // static HMC instead of NUTS (the hmc_static_* services: base_static_hmc.hpp:41-70)
  double int_time = 6.28318530717958647692;
};

// metric_writer interface: one JSON record per chain (callbacks::structured_writer's use in write_sampler_state_struct)
class json_metric_writer {
 public:
  json_metric_writer() {}
  explicit json_metric_writer(std::unique_ptr<std::ostream>&& out) : out_(std::move(out)) {}
  void write(double stepsize, const double* inv_metric, int64_t n, bool dense = false) {
    if (!out_) return;
    *out_ << "{\n  \"stepsize\" : " << stepsize << ",\n  \"metric_type\" : \"" << (dense ? "dense_e" : "diag_e") << "\",\n  \"inv_metric\" : [ ";
    if (dense) {  // json_writer::write(key, Eigen::MatrixXd), callbacks/json_writer.hpp:484-500: one bracketed row per matrix row
      for (int64_t i = 0; i < n; i++) {
        *out_ << (i ? ", " : "") << "[ ";
        for (int64_t j = 0; j < n; j++) *out_ << (j ? ", " : "") << inv_metric[i * n + j];
        *out_ << " ]";
      }
    } else {
      for (int64_t i = 0; i < n; i++) *out_ << (i ? ", " : "") << inv_metric[i];
    }
    *out_ << " ]\n}\n";
    out_->flush();
  }

 private:
  std::unique_ptr<std::ostream> out_;
};

namespace sample {

namespace internal {

struct engine_guard {
  std::vector<b200nuts_engine*> e;
  ~engine_guard() {
    for (auto p : e) b200nuts_destroy(p);
  }
};

// one CSV row at the stream's precision; identical to `os << v[0] << "," << ...` of unique_stream_writer::write_vector
inline void format_row(std::string& buf, const double* v, int64_t n, int precision) {
  char tmp[40];
  for (int64_t i = 0; i < n; i++) {
    int len = std::snprintf(tmp, sizeof tmp, "%.*g", precision, v[i]);
    buf.append(tmp, (size_t)len);
    buf.push_back(i + 1 < n ? ',' : '\n');
  }
}

// iterations (0-based, within a phase of `n` transitions starting at global index `start` of `finish`) at which
// generate_transitions logs a progress message
inline std::vector<int> message_iterations(int n, int start, int finish, int refresh) {
  std::vector<int> r;
  if (refresh <= 0) return r;
  for (int m = 0; m < n; ++m)
    if (start + m + 1 == finish || m == 0 || (m + 1) % refresh == 0) r.push_back(m);
  return r;
}

inline std::string progress_message(int m, int start, int finish, bool warmup, size_t chain_id, size_t num_chains) {
  int it_print_width = (int)std::ceil(std::log10(static_cast<double>(finish)));
  std::stringstream message;
  if (num_chains != 1) message << "Chain [" << chain_id << "] ";
  message << "Iteration: ";
  message << std::setw(it_print_width) << m + 1 + start << " / " << finish;
  message << " [" << std::setw(3) << static_cast<int>((100.0 * (start + m + 1)) / finish) << "%] ";
  message << (warmup ? " (Warmup)" : " (Sampling)");
  return message.str();
}

}  // namespace internal

template <class Model, typename InitContextPtr, typename InitInvContextPtr, typename InitWriter, typename SampleWriter,
          typename DiagnosticWriter, typename MetricWriter>
int hmc_nuts_diag_e_adapt(Model& model, size_t num_chains, const std::vector<InitContextPtr>& init,
                          const std::vector<InitInvContextPtr>& init_inv_metric, unsigned int random_seed,
                          unsigned int init_chain_id, double init_radius, int num_warmup, int num_samples, int num_thin,
                          bool save_warmup, int refresh, double stepsize, double stepsize_jitter, int max_depth, double delta,
                          double gamma, double kappa, double t0, unsigned int init_buffer, unsigned int term_buffer,
                          unsigned int window, callbacks::interrupt& interrupt, callbacks::logger& logger,
                          std::vector<InitWriter>& init_writer, std::vector<SampleWriter>& sample_writer,
                          std::vector<DiagnosticWriter>& diagnostic_writer, std::vector<MetricWriter>& metric_writer,
                          const device_options& opt = device_options(), bool adapt_engaged = true) {
  using internal::engine_guard;
  (void)init_writer;
  const int64_t D = model.num_params_r();
  const int C = (int)num_chains;
  // ---- inits and metric (initialize.hpp:70-260; util::read_diag_inv_metric / validate_diag_inv_metric)
  std::vector<double> q0, im;
  try {
    size_t n_user = 0;
    for (size_t c = 0; c < num_chains && c < init.size(); c++)
      if (init[c] && !init[c]->empty()) n_user++;
    if (n_user != 0 && n_user != num_chains) throw std::invalid_argument("inits must be given for every chain or for none");
    if (n_user) {
      std::vector<double> one;
      for (size_t c = 0; c < num_chains; c++) {
        model.transform_inits(*init[c], one);
        q0.insert(q0.end(), one.begin(), one.end());
      }
    }
    bool any_metric = false;
    for (size_t c = 0; c < num_chains && c < init_inv_metric.size(); c++)
      if (init_inv_metric[c] && init_inv_metric[c]->contains_r("inv_metric")) any_metric = true;
    if (any_metric) {
      // util::read_diag_inv_metric / read_dense_inv_metric (services/util/read_*_inv_metric.hpp); a dense matrix is symmetric,
      // so the flattening order of the context does not matter
      const int64_t MD = opt.dense_e ? D * D : D;
      im.assign((size_t)C * MD, opt.dense_e ? 0.0 : 1.0);
      if (opt.dense_e)
        for (int c = 0; c < C; c++)
          for (int64_t i = 0; i < D; i++) im[(size_t)c * MD + i * D + i] = 1.0;
      for (size_t c = 0; c < num_chains && c < init_inv_metric.size(); c++) {
        if (!init_inv_metric[c] || !init_inv_metric[c]->contains_r("inv_metric")) continue;
        const std::vector<double>& v = init_inv_metric[c]->vals_r("inv_metric");
        if ((int64_t)v.size() != MD) throw std::domain_error("Cannot get inverse metric from input file; expected " + std::to_string(MD) + " elements");
        std::copy(v.begin(), v.end(), im.begin() + c * MD);
      }
    }
  } catch (const std::domain_error& e) {
    logger.error(e.what());
    return error_codes::CONFIG;
  } catch (const std::exception& e) {
    logger.error(e.what());
    return error_codes::CONFIG;
  }

  // a diagnostic writer with a stream wants the momentum and gradient of every saved draw (mcmc_writer.hpp:156-176)
  bool want_diag = false;
  for (auto& w : diagnostic_writer)
    if (w.stream()) want_diag = true;

  // ---- engines: chains split contiguously over the devices
  const int G = (int)std::max<size_t>(1, std::min<size_t>(opt.devices.size(), num_chains));
  engine_guard eg;
  std::vector<int> first(G + 1, 0);
  for (int g = 0; g < G; g++) first[g + 1] = (int)((int64_t)C * (g + 1) / G);
  for (int g = 0; g < G; g++) {
    b200nuts_config cfg;
    b200nuts_default_config(&cfg);
    cfg.model = model.model_id();
    cfg.num_chains = first[g + 1] - first[g];
    cfg.random_seed = random_seed;
    cfg.init_chain_id = init_chain_id + (unsigned)first[g];
    cfg.init_radius = init_radius;
    cfg.num_warmup = num_warmup;
    cfg.num_samples = num_samples;
    cfg.num_thin = num_thin;
    cfg.save_warmup = save_warmup;
    cfg.stepsize = stepsize;
    cfg.stepsize_jitter = stepsize_jitter;
    cfg.max_depth = max_depth;
    cfg.delta = delta;
    cfg.gamma = gamma;
    cfg.kappa = kappa;
    cfg.t0 = t0;
    cfg.init_buffer = init_buffer;
    cfg.term_buffer = term_buffer;
    cfg.window = window;
    cfg.adapt_engaged = adapt_engaged;
    cfg.metric = opt.dense_e ? 2 : (opt.unit_e ? 1 : 0);
    cfg.store_diagnostics = want_diag ? 1 : 0;
    cfg.static_hmc = opt.static_hmc ? 1 : 0;
    cfg.int_time = opt.int_time;
    cfg.engine = opt.engine;
    cfg.device = opt.devices.empty() ? 0 : opt.devices[g];
    b200nuts_engine* e = nullptr;
    int rc = b200nuts_create(&cfg, &e);
    if (rc) {
      logger.error(b200nuts_last_error(nullptr));
      return rc;
    }
    eg.e.push_back(e);
    if ((rc = model.upload(e)) || (!q0.empty() && (rc = b200nuts_set_inits(e, q0.data() + (size_t)first[g] * D))) ||
        (!im.empty() && (rc = b200nuts_set_inv_metric(e, im.data() + (size_t)first[g] * D * (opt.dense_e ? D : 1)))) ||
        (rc = b200nuts_init_chains(e))) {
      logger.error(b200nuts_last_error(e));
      return rc;
    }
  }
  int64_t nu = 0, K = 0;
  b200nuts_num_params(eg.e[0], &nu, &K);

  auto run_all = [&](int n_iter) -> int {
    std::vector<int> rcs(G, 0);
    std::vector<std::thread> th;
    for (int g = 1; g < G; g++) th.emplace_back([&, g] { rcs[g] = b200nuts_run(eg.e[g], n_iter); });
    rcs[0] = b200nuts_run(eg.e[0], n_iter);
    for (auto& t : th) t.join();
    for (int g = 0; g < G; g++)
      if (rcs[g]) {
        logger.error(b200nuts_last_error(eg.e[g]));
        return rcs[g];
      }
    return 0;
  };
  // one phase of `n` transitions with the reference's progress messages; the interrupt runs at each message point
  auto run_phase = [&](int n, int start, bool warmup) -> int {
    const int finish = num_warmup + num_samples;
    std::vector<int> at = internal::message_iterations(n, start, finish, refresh);
    int done = 0;
    size_t k = 0;
    while (done < n) {
      if (k < at.size() && at[k] == done) {
        interrupt();
        if (C <= 16) {
          for (int c = 0; c < C; c++) logger.info(internal::progress_message(done, start, finish, warmup, init_chain_id + c, num_chains));
        } else {
          std::stringstream m;
          m << "Chains [" << init_chain_id << "-" << init_chain_id + C - 1 << "] "
            << internal::progress_message(done, start, finish, warmup, 0, 1);
          logger.info(m.str());
        }
        k++;
      }
      const int next = k < at.size() ? at[k] : n;
      int rc = run_all(next - done);
      if (rc) return rc;
      done = next;
    }
    return 0;
  };
  // rows [row0, row0 + nrows) of every chain -> its sample writer
  auto write_rows = [&](int64_t row0, int64_t nrows) -> int {
    if (nrows <= 0) return 0;
    const int64_t ncol = 7 + K;
    for (int g = 0; g < G; g++) {
      const int Cg = first[g + 1] - first[g];
      // bound the host staging buffer to ~1 GiB
      int64_t chunk = std::max<int64_t>(1, (int64_t)((1ull << 30) / (sizeof(double) * (size_t)ncol * (size_t)nrows)));
      std::vector<double> buf;
      for (int64_t c0 = 0; c0 < Cg; c0 += chunk) {
        const int64_t m = std::min<int64_t>(chunk, Cg - c0);
        buf.resize((size_t)m * nrows * ncol);
        int rc = b200nuts_fetch_draws(eg.e[g], c0, m, row0, nrows, 1, buf.data());
        if (rc) {
          logger.error(b200nuts_last_error(eg.e[g]));
          return rc;
        }
        const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(opt.format_threads, m));
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++)
          th.emplace_back([&, t] {
            std::string text;
            for (int64_t c = t; c < m; c += nt) {
              const size_t chain = (size_t)(first[g] + c0 + c);
              if (chain >= sample_writer.size()) continue;
              std::ostream* os = sample_writer[chain].stream();
              if (!os) continue;
              text.clear();
              const int prec = (int)os->precision();
              for (int64_t r = 0; r < nrows; r++) {
                double* row = buf.data() + ((size_t)c * nrows + r) * ncol;
                if (opt.static_hmc) {  // engine row: lp, accept, stepsize, L, 1, 0, energy, values -> lp, accept, stepsize, int_time, energy, values
                  row[4] = row[2];
                  row[3] = row[1];
                  row[2] = row[0];
                  row[5] = opt.int_time;
                  internal::format_row(text, row + 2, ncol - 2, prec);
                } else {
                  internal::format_row(text, row, ncol, prec);
                }
              }
              os->write(text.data(), (std::streamsize)text.size());
            }
          });
        for (auto& t : th) t.join();
      }
    }
    return 0;
  };
  auto chain_errors = [&]() -> int {
    int code = 0;
    for (int g = 0; g < G; g++) {
      const int Cg = first[g + 1] - first[g];
      std::vector<int32_t> err(Cg);
      if (b200nuts_get_chain_status(eg.e[g], err.data(), nullptr, nullptr)) return error_codes::SOFTWARE;
      for (int c = 0; c < Cg; c++)
        if (err[c]) {
          std::stringstream m;
          if (num_chains != 1) m << "Chain [" << init_chain_id + first[g] + c << "] ";
          m << (err[c] == error_codes::CONFIG ? "Initialization failed." : "Exception: step size or metric adaptation failed (Posterior is improper / No acceptably small step size / Numerical overflow in metric adaptation).");
          logger.error(m.str());
          if (!code) code = err[c];
        }
    }
    return code;
  };

  const int64_t warm_rows = save_warmup ? (num_warmup + num_thin - 1) / num_thin : 0;
  const int64_t samp_rows = (num_samples + num_thin - 1) / num_thin;
  int rc;
  // ---- util::initialize + sampler set-up of every chain come first in the reference (hmc_nuts_diag_e_adapt.hpp:360-379): a
  // failed initialisation ends the call before any notice, header row or iteration message
  if ((rc = run_all(0))) return rc;
  if ((rc = chain_errors())) return rc;

  // window fallback notice (windowed_adaptation.hpp:42-67), once -- the reference prints it per chain
  if (adapt_engaged && !opt.unit_e && num_warmup >= 20 && init_buffer + window + term_buffer > (unsigned)num_warmup) {
    unsigned ib = (unsigned)(0.15 * num_warmup), tb = (unsigned)(0.1 * num_warmup);
    std::stringstream m;
    m << "WARNING: There aren't enough warmup iterations to fit the\n         three stages of adaptation as currently configured.\n"
      << "         Reducing each adaptation stage to 15%/75%/10% of\n         the given number of warmup iterations:\n"
      << "           init_buffer = " << ib << "\n           adapt_window = " << num_warmup - (ib + tb) << "\n           term_buffer = " << tb << "\n";
    logger.info(m.str());
  } else if (adapt_engaged && !opt.unit_e && num_warmup < 20) {
    // windowed_adaptation's constructor names the estimator: "variance" (var_adaptation.hpp), "covariance" (covar_adaptation.hpp)
    logger.info(std::string("WARNING: No ") + (opt.dense_e ? "covariance" : "variance") + " estimation is\n         performed for num_warmup < 20\n");
  }

  // ---- header rows
  std::vector<std::string> names = {"lp__", "accept_stat__", "stepsize__", "treedepth__", "n_leapfrog__", "divergent__", "energy__"};
  if (opt.static_hmc) names = {"lp__", "accept_stat__", "stepsize__", "int_time__", "energy__"};  // base_static_hmc::get_sampler_param_names
  model.constrained_param_names(names, true, true);
  // diagnostic header (base_hmc::get_sampler_diagnostic_names -> ps_point::get_param_names): the unconstrained parameter names,
  // then p_<name>, then g_<name>
  std::vector<std::string> diag_names(names.begin(), names.begin() + 7);
  {
    std::vector<std::string> un;
    model.unconstrained_param_names(un, false, false);
    for (const std::string& n : un) diag_names.push_back(n);
    for (const std::string& n : un) diag_names.push_back("p_" + n);
    for (const std::string& n : un) diag_names.push_back("g_" + n);
  }
  for (int c = 0; c < C; c++) {
    if ((size_t)c < sample_writer.size()) sample_writer[c](names);
    if ((size_t)c < diagnostic_writer.size()) diagnostic_writer[c](diag_names);
  }
  // rows [row0, row0 + nrows) of every chain -> its diagnostic writer (sampler columns, q, p, g)
  auto write_diag_rows = [&](int64_t row0, int64_t nrows) -> int {
    if (!want_diag || nrows <= 0) return 0;
    const int64_t ncol = 7 + 3 * D;
    std::vector<double> buf;
    std::string text;
    for (int g = 0; g < G; g++) {
      const int Cg = first[g + 1] - first[g];
      int64_t chunk = std::max<int64_t>(1, (int64_t)((1ull << 30) / (sizeof(double) * (size_t)ncol * (size_t)nrows)));
      for (int64_t c0 = 0; c0 < Cg; c0 += chunk) {
        const int64_t m = std::min<int64_t>(chunk, Cg - c0);
        buf.resize((size_t)m * nrows * ncol);
        int rc = b200nuts_fetch_diagnostics(eg.e[g], c0, m, row0, nrows, buf.data());
        if (rc) {
          logger.error(b200nuts_last_error(eg.e[g]));
          return rc;
        }
        for (int64_t c = 0; c < m; c++) {
          const size_t chain = (size_t)(first[g] + c0 + c);
          if (chain >= diagnostic_writer.size()) continue;
          std::ostream* os = diagnostic_writer[chain].stream();
          if (!os) continue;
          text.clear();
          const int prec = (int)os->precision();
          for (int64_t r = 0; r < nrows; r++) internal::format_row(text, buf.data() + ((size_t)c * nrows + r) * ncol, ncol, prec);
          os->write(text.data(), (std::streamsize)text.size());
        }
      }
    }
    return 0;
  };

  // ---- warm-up (generate_transitions with warmup = true)
  auto t0w = std::chrono::steady_clock::now();
  if ((rc = run_phase(num_warmup, 0, true))) return rc;
  double warm_delta_t = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0w).count() / 1000.0;
  if ((rc = chain_errors())) return rc;
  if ((rc = write_rows(0, warm_rows))) return rc;
  if ((rc = write_diag_rows(0, warm_rows))) return rc;
  // ---- adaptation block (write_adapt_finish + write_sampler_state) and metric JSON
  {
    const int64_t MD = opt.dense_e ? D * D : D;
    std::vector<double> eps((size_t)C), met((size_t)C * MD);
    for (int g = 0; g < G; g++)
      if ((rc = b200nuts_get_adaptation(eg.e[g], eps.data() + first[g], met.data() + (size_t)first[g] * MD))) {
        logger.error(b200nuts_last_error(eg.e[g]));
        return rc;
      }
    for (int c = 0; c < C; c++) {
      if ((size_t)c < sample_writer.size()) {
        auto& w = sample_writer[c];
        {  // util::run_sampler (no adaptation) writes the same block: run_sampler.hpp:66-67, run_adaptive_sampler.hpp:88-89
          w("Adaptation terminated");
          std::stringstream ss;
          ss << "Step size = " << eps[c];
          w(ss.str());
          if (opt.unit_e) {
            w("No free parameters for unit metric");  // unit_e_point::write_metric (mcmc/hmc/hamiltonians/unit_e_point.hpp)
          } else if (opt.dense_e) {  // dense_e_point::write_metric (mcmc/hmc/hamiltonians/dense_e_point.hpp:46-56): one row per line
            w("Elements of inverse mass matrix:");
            for (int64_t i = 0; i < D; i++) {
              std::stringstream ms;
              for (int64_t j = 0; j < D; j++) ms << (j ? ", " : "") << met[(size_t)c * MD + i * D + j];
              w(ms.str());
            }
          } else {
            w("Diagonal elements of inverse mass matrix:");
            std::stringstream ms;
            for (int64_t i = 0; i < D; i++) ms << (i ? ", " : "") << met[(size_t)c * D + i];
            w(ms.str());
          }
        }
      }
      if (adapt_engaged && (size_t)c < metric_writer.size()) metric_writer[c].write(eps[c], met.data() + (size_t)c * MD, D, opt.dense_e);
    }
  }
  // ---- sampling
  auto t0s = std::chrono::steady_clock::now();
  if ((rc = run_phase(num_samples, num_warmup, false))) return rc;
  double sample_delta_t = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0s).count() / 1000.0;
  if ((rc = chain_errors())) return rc;
  if ((rc = write_rows(warm_rows, samp_rows))) return rc;
  if ((rc = write_diag_rows(warm_rows, samp_rows))) return rc;
  // ---- timing footer (mcmc_writer::write_timing)
  auto timing = [&](auto& w) {
    const std::string title(" Elapsed Time: ");
    w();
    std::stringstream s1, s2, s3;
    s1 << title << warm_delta_t << " seconds (Warm-up)";
    s2 << std::string(title.size(), ' ') << sample_delta_t << " seconds (Sampling)";
    s3 << std::string(title.size(), ' ') << warm_delta_t + sample_delta_t << " seconds (Total)";
    w(s1.str());
    w(s2.str());
    w(s3.str());
    w();
  };
  for (int c = 0; c < C; c++) {
    if ((size_t)c < sample_writer.size()) timing(sample_writer[c]);
    if ((size_t)c < diagnostic_writer.size()) timing(diagnostic_writer[c]);
  }
  {
    const std::string title(" Elapsed Time: ");
    std::stringstream s1, s2, s3;
    s1 << title << warm_delta_t << " seconds (Warm-up)";
    s2 << std::string(title.size(), ' ') << sample_delta_t << " seconds (Sampling)";
    s3 << std::string(title.size(), ' ') << warm_delta_t + sample_delta_t << " seconds (Total)";
    logger.info("");
    logger.info(s1.str());
    logger.info(s2.str());
    logger.info(s3.str());
    logger.info("");
  }
  return error_codes::OK;
}

// ---- sibling services: same engine, fewer knobs; parameter lists as in the reference ----------

// stan::services::sample::hmc_nuts_diag_e (services/sample/hmc_nuts_diag_e.hpp): no adaptation
template <class Model, typename InitContextPtr, typename InitInvContextPtr, typename InitWriter, typename SampleWriter,
          typename DiagnosticWriter>
int hmc_nuts_diag_e(Model& model, size_t num_chains, const std::vector<InitContextPtr>& init,
                    const std::vector<InitInvContextPtr>& init_inv_metric, unsigned int random_seed, unsigned int init_chain_id,
                    double init_radius, int num_warmup, int num_samples, int num_thin, bool save_warmup, int refresh, double stepsize,
                    double stepsize_jitter, int max_depth, callbacks::interrupt& interrupt, callbacks::logger& logger,
                    std::vector<InitWriter>& init_writer, std::vector<SampleWriter>& sample_writer,
                    std::vector<DiagnosticWriter>& diagnostic_writer, const device_options& opt = device_options()) {
  std::vector<json_metric_writer> none;
  return hmc_nuts_diag_e_adapt(model, num_chains, init, init_inv_metric, random_seed, init_chain_id, init_radius, num_warmup, num_samples,
                               num_thin, save_warmup, refresh, stepsize, stepsize_jitter, max_depth, 0.8, 0.05, 0.75, 10.0, 75u, 50u, 25u,
                               interrupt, logger, init_writer, sample_writer, diagnostic_writer, none, opt, false);
}

// stan::services::sample::hmc_nuts_dense_e_adapt (services/sample/hmc_nuts_dense_e_adapt.hpp:331-407 multi-chain): dense
// Euclidean metric; step size and covariance adapted in windows (adapt_dense_e_nuts.hpp:25-44, covar_adaptation.hpp:17-45).
// init_inv_metric contexts hold "inv_metric" as a D x D matrix (util::read_dense_inv_metric).
template <class Model, typename InitContextPtr, typename InitInvContextPtr, typename InitWriter, typename SampleWriter,
          typename DiagnosticWriter, typename MetricWriter>
int hmc_nuts_dense_e_adapt(Model& model, size_t num_chains, const std::vector<InitContextPtr>& init,
                           const std::vector<InitInvContextPtr>& init_inv_metric, unsigned int random_seed, unsigned int init_chain_id,
                           double init_radius, int num_warmup, int num_samples, int num_thin, bool save_warmup, int refresh,
                           double stepsize, double stepsize_jitter, int max_depth, double delta, double gamma, double kappa, double t0,
                           unsigned int init_buffer, unsigned int term_buffer, unsigned int window, callbacks::interrupt& interrupt,
                           callbacks::logger& logger, std::vector<InitWriter>& init_writer, std::vector<SampleWriter>& sample_writer,
                           std::vector<DiagnosticWriter>& diagnostic_writer, std::vector<MetricWriter>& metric_writer,
                           device_options opt = device_options(), bool adapt_engaged = true) {
  opt.dense_e = true;
  opt.unit_e = false;
  return hmc_nuts_diag_e_adapt(model, num_chains, init, init_inv_metric, random_seed, init_chain_id, init_radius, num_warmup, num_samples,
                               num_thin, save_warmup, refresh, stepsize, stepsize_jitter, max_depth, delta, gamma, kappa, t0, init_buffer,
                               term_buffer, window, interrupt, logger, init_writer, sample_writer, diagnostic_writer, metric_writer, opt,
                               adapt_engaged);
}

// stan::services::sample::hmc_nuts_dense_e (services/sample/hmc_nuts_dense_e.hpp): dense metric, no adaptation
template <class Model, typename InitContextPtr, typename InitInvContextPtr, typename InitWriter, typename SampleWriter,
          typename DiagnosticWriter>
int hmc_nuts_dense_e(Model& model, size_t num_chains, const std::vector<InitContextPtr>& init,
                     const std::vector<InitInvContextPtr>& init_inv_metric, unsigned int random_seed, unsigned int init_chain_id,
                     double init_radius, int num_warmup, int num_samples, int num_thin, bool save_warmup, int refresh, double stepsize,
                     double stepsize_jitter, int max_depth, callbacks::interrupt& interrupt, callbacks::logger& logger,
                     std::vector<InitWriter>& init_writer, std::vector<SampleWriter>& sample_writer,
                     std::vector<DiagnosticWriter>& diagnostic_writer, const device_options& opt = device_options()) {
  std::vector<json_metric_writer> none;
  return hmc_nuts_dense_e_adapt(model, num_chains, init, init_inv_metric, random_seed, init_chain_id, init_radius, num_warmup, num_samples,
                                num_thin, save_warmup, refresh, stepsize, stepsize_jitter, max_depth, 0.8, 0.05, 0.75, 10.0, 75u, 50u, 25u,
                                interrupt, logger, init_writer, sample_writer, diagnostic_writer, none, opt, false);
}

// ---- static HMC (services/sample/hmc_static_{diag,unit,dense}_e[_adapt].hpp; base_static_hmc.hpp:41-70): same engine, L =
// max(1, int(int_time / stepsize)) leapfrogs and a Metropolis test per transition; columns lp__, accept_stat__, stepsize__,
// int_time__, energy__.  Parameter lists as in the reference (int_time where the NUTS services have max_depth).
#define B200_STATIC_SERVICE(NAME, DENSE, UNIT, ADAPT)                                                                                        \
  template <class Model, typename InitContextPtr, typename InitInvContextPtr, typename InitWriter, typename SampleWriter, typename DiagnosticWriter> \
  int NAME(Model& model, size_t num_chains, const std::vector<InitContextPtr>& init, const std::vector<InitInvContextPtr>& init_inv_metric,   \
           unsigned int random_seed, unsigned int init_chain_id, double init_radius, int num_warmup, int num_samples, int num_thin,         \
           bool save_warmup, int refresh, double stepsize, double stepsize_jitter, double int_time, double delta, double gamma,             \
           double kappa, double t0, unsigned int init_buffer, unsigned int term_buffer, unsigned int window,                                \
           callbacks::interrupt& interrupt, callbacks::logger& logger, std::vector<InitWriter>& init_writer,                                \
           std::vector<SampleWriter>& sample_writer, std::vector<DiagnosticWriter>& diagnostic_writer, device_options opt = device_options()) { \
    opt.static_hmc = true;                                                                                                                   \
    opt.int_time = int_time;                                                                                                                 \
    opt.dense_e = DENSE;                                                                                                                     \
    opt.unit_e = UNIT;                                                                                                                       \
    std::vector<json_metric_writer> none;                                                                                                    \
    return hmc_nuts_diag_e_adapt(model, num_chains, init, init_inv_metric, random_seed, init_chain_id, init_radius, num_warmup, num_samples, \
                                 num_thin, save_warmup, refresh, stepsize, stepsize_jitter, 10, delta, gamma, kappa, t0, init_buffer,        \
                                 term_buffer, window, interrupt, logger, init_writer, sample_writer, diagnostic_writer, none, opt, ADAPT);   \
  }
B200_STATIC_SERVICE(hmc_static_diag_e_adapt, false, false, true)
B200_STATIC_SERVICE(hmc_static_unit_e_adapt, false, true, true)
B200_STATIC_SERVICE(hmc_static_dense_e_adapt, true, false, true)
// the non-adaptive services take no adaptation parameters (services/sample/hmc_static_diag_e.hpp:62-70)
#define B200_STATIC_SERVICE_FIXED(NAME, DENSE, UNIT)                                                                                         \
  template <class Model, typename InitContextPtr, typename InitInvContextPtr, typename InitWriter, typename SampleWriter, typename DiagnosticWriter> \
  int NAME(Model& model, size_t num_chains, const std::vector<InitContextPtr>& init, const std::vector<InitInvContextPtr>& init_inv_metric,   \
           unsigned int random_seed, unsigned int init_chain_id, double init_radius, int num_warmup, int num_samples, int num_thin,         \
           bool save_warmup, int refresh, double stepsize, double stepsize_jitter, double int_time, callbacks::interrupt& interrupt,         \
           callbacks::logger& logger, std::vector<InitWriter>& init_writer, std::vector<SampleWriter>& sample_writer,                       \
           std::vector<DiagnosticWriter>& diagnostic_writer, device_options opt = device_options()) {                                       \
    opt.static_hmc = true;                                                                                                                   \
    opt.int_time = int_time;                                                                                                                 \
    opt.dense_e = DENSE;                                                                                                                     \
    opt.unit_e = UNIT;                                                                                                                       \
    std::vector<json_metric_writer> none;                                                                                                    \
    return hmc_nuts_diag_e_adapt(model, num_chains, init, init_inv_metric, random_seed, init_chain_id, init_radius, num_warmup, num_samples, \
                                 num_thin, save_warmup, refresh, stepsize, stepsize_jitter, 10, 0.8, 0.05, 0.75, 10.0, 75u, 50u, 25u,        \
                                 interrupt, logger, init_writer, sample_writer, diagnostic_writer, none, opt, false);                        \
  }
B200_STATIC_SERVICE_FIXED(hmc_static_diag_e, false, false)
B200_STATIC_SERVICE_FIXED(hmc_static_unit_e, false, true)
B200_STATIC_SERVICE_FIXED(hmc_static_dense_e, true, false)
#undef B200_STATIC_SERVICE_FIXED
#undef B200_STATIC_SERVICE

// stan::services::sample::hmc_nuts_unit_e_adapt (services/sample/hmc_nuts_unit_e_adapt.hpp): identity metric, step size adapted
template <class Model, typename InitContextPtr, typename InitWriter, typename SampleWriter, typename DiagnosticWriter,
          typename MetricWriter>
int hmc_nuts_unit_e_adapt(Model& model, size_t num_chains, const std::vector<InitContextPtr>& init, unsigned int random_seed,
                          unsigned int init_chain_id, double init_radius, int num_warmup, int num_samples, int num_thin,
                          bool save_warmup, int refresh, double stepsize, double stepsize_jitter, int max_depth, double delta,
                          double gamma, double kappa, double t0, callbacks::interrupt& interrupt, callbacks::logger& logger,
                          std::vector<InitWriter>& init_writer, std::vector<SampleWriter>& sample_writer,
                          std::vector<DiagnosticWriter>& diagnostic_writer, std::vector<MetricWriter>& metric_writer,
                          device_options opt = device_options(), bool adapt_engaged = true) {
  opt.unit_e = true;
  std::vector<std::shared_ptr<io::var_context>> no_metric;
  return hmc_nuts_diag_e_adapt(model, num_chains, init, no_metric, random_seed, init_chain_id, init_radius, num_warmup, num_samples, num_thin,
                               save_warmup, refresh, stepsize, stepsize_jitter, max_depth, delta, gamma, kappa, t0, 75u, 50u, 25u, interrupt,
                               logger, init_writer, sample_writer, diagnostic_writer, metric_writer, opt, adapt_engaged);
}

// stan::services::sample::hmc_nuts_unit_e (services/sample/hmc_nuts_unit_e.hpp): identity metric, no adaptation
template <class Model, typename InitContextPtr, typename InitWriter, typename SampleWriter, typename DiagnosticWriter>
int hmc_nuts_unit_e(Model& model, size_t num_chains, const std::vector<InitContextPtr>& init, unsigned int random_seed,
                    unsigned int init_chain_id, double init_radius, int num_warmup, int num_samples, int num_thin, bool save_warmup,
                    int refresh, double stepsize, double stepsize_jitter, int max_depth, callbacks::interrupt& interrupt,
                    callbacks::logger& logger, std::vector<InitWriter>& init_writer, std::vector<SampleWriter>& sample_writer,
                    std::vector<DiagnosticWriter>& diagnostic_writer, const device_options& opt = device_options()) {
  std::vector<json_metric_writer> none;
  return hmc_nuts_unit_e_adapt(model, num_chains, init, random_seed, init_chain_id, init_radius, num_warmup, num_samples, num_thin,
                               save_warmup, refresh, stepsize, stepsize_jitter, max_depth, 0.8, 0.05, 0.75, 10.0, interrupt, logger,
                               init_writer, sample_writer, diagnostic_writer, none, opt, false);
}

}  // namespace sample
}  // namespace services
}  // namespace b200
