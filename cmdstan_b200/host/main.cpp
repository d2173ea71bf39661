// b200stan -- the `model sample ...` front end of CmdStan (src/cmdstan/main.cpp, cmdstan::command() in
// src/cmdstan/command.hpp:84-760, `sample` branch :560-700) on top of the B200 NUTS service.
//
//   <model> sample [num_samples=..] [num_warmup=..] [save_warmup=..] [thin=..] [adapt ...] [algorithm=hmc [engine=nuts
//           [max_depth=..]] [metric=diag_e] [metric_file=..] [stepsize=..] [stepsize_jitter=..]] [num_chains=..]
//           [id=..] [data file=..] [init=..] [random seed=..] [output file=.. refresh=.. sig_figs=..] [num_threads=..]
//
// The model is selected by the executable's name (one executable per model, as in CmdStan: `bernoulli sample ...`;
// hard links / symlinks to this binary) or, when invoked as `b200stan`, by the first argument.
// Output: one Stan CSV per chain (`<base>_<id+i>.csv` when num_chains > 1, file.hpp:189-212) with the reference's
// header (write_stan / write_model / write_datetime / argument echo / compile info, command_helper.hpp), readable by
// the reference's stansummary / diagnose.  GPUs: env B200_DEVICES="0,1,..." (default "0"); chains are split across them.
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <fstream>
#include <iostream>
#include <random>

#include "arguments.hpp"
#include "service.hpp"

namespace {

std::pair<std::string, std::string> basename_suffix(const std::string& name) {  // file.hpp:92-106
  size_t slash = name.find_last_of("/\\");
  size_t dot = name.find_last_of('.');
  if (dot == std::string::npos || (slash != std::string::npos && dot < slash) || name.size() - dot <= 1) return {name, ""};
  return {name.substr(0, dot), name.substr(dot)};
}

std::vector<std::string> make_filenames(const std::string& filename, const std::string& tag, const std::string& type, unsigned num_chains,
                                        unsigned id) {  // file.hpp:189-212
  auto bs = basename_suffix(filename);
  if (type != ".csv" || bs.second.empty()) bs.second = type;
  std::vector<std::string> names(num_chains);
  for (unsigned i = 0; i < num_chains; ++i) names[i] = bs.first + tag + (num_chains == 1 ? std::string("") : "_" + std::to_string(i + id)) + bs.second;
  return names;
}

std::unique_ptr<std::ostream> create_stream(const std::string& path, int sig_figs) {
  std::unique_ptr<std::ofstream> f(new std::ofstream(path.c_str()));
  if (!f->good()) throw std::invalid_argument("Cannot write to file: " + path);
  if (sig_figs > -1) f->precision(sig_figs);
  return std::unique_ptr<std::ostream>(std::move(f));
}

std::string model_from_argv0(const std::string& a0) {
  size_t slash = a0.find_last_of('/');
  return slash == std::string::npos ? a0 : a0.substr(slash + 1);
}

// method=log_prob (src/cmdstan/command_helper.hpp:520-600 services_log_prob_grad): log density and its gradient at sets of
// parameter values read from JSON -- `unconstrained_params` {"params_r": [..] or [[..], ..]} or `constrained_params` (one set of
// named values, mapped through transform_inits) -- evaluated on the device (b200nuts_log_prob_grad), written as "lp__, g_<name>.."
int run_log_prob(b200::cmd::argument_parser& parser, const std::string& model_name) {
  using namespace b200;
  const std::string upath = parser.str({"method", "log_prob", "unconstrained_params"});
  const std::string cpath = parser.str({"method", "log_prob", "constrained_params"});
  const bool jacobian = parser.boolean({"method", "log_prob", "jacobian"});
  const int sig_figs = (int)parser.integer({"output", "sig_figs"});
  if (upath.empty() == cpath.empty()) {
    std::cerr << (upath.empty() ? "No input parameter files provided, call to log_prob done." : "Cannot specify both input files of both constrained and unconstrained parameter values.") << std::endl;
    return services::error_codes::USAGE;
  }
  long long seed_arg = parser.integer({"random", "seed"});
  if (seed_arg < 0) {
    const unsigned int s = (unsigned int)(std::time(nullptr) % 4294967296ULL);
    parser.find({"random", "seed"})->value = parser.find({"random", "seed"})->def = std::to_string(s);
  }
  for (const std::string& ln : parser.echo()) std::cout << ln << std::endl;
  std::cout << std::endl;
  b200nuts_engine* e = nullptr;
  try {
    io::var_context data = io::read_json_file(parser.str({"data", "file"}));
    model::device_model& model = model::new_model(model_name, data, 0, &std::cout);
    const int64_t D = model.num_params_r();
    std::vector<double> q;
    if (!upath.empty()) {
      io::var_context u = io::read_json_file(upath);
      const std::vector<double>& v = u.vals_r("params_r");
      const std::vector<int64_t>& dims = u.dims_r("params_r");
      if (dims.size() <= 1) {
        if ((int64_t)v.size() != D) throw std::invalid_argument("Incorrect number of unconstrained parameters provided! Model has " + std::to_string(D) + " parameters but " + std::to_string(v.size()) + " were found.");
        q = v;
      } else {
        const int64_t n = dims[0], d = dims[1];
        if (d != D) throw std::invalid_argument("Incorrect number of unconstrained parameters provided! Model has " + std::to_string(D) + " parameters but " + std::to_string(d) + " were found.");
        // the reference's JSON reader flattens the matrix column-major and get_uparams_r re-reads it row-major
        // (command_helper.hpp:548-556): element (i, j) of set i is the flat element i*D + j of that column-major buffer
        q.resize((size_t)n * D);
        for (int64_t k = 0; k < n * D; k++) q[(size_t)k] = v[(size_t)((k % n) * D + k / n)];
      }
    } else {
      io::var_context c = io::read_json_file(cpath);
      model.transform_inits(c, q);
    }
    const int64_t n = (int64_t)q.size() / D;
    b200nuts_config cfg;
    b200nuts_default_config(&cfg);
    cfg.model = model.model_id();
    int rc = b200nuts_create(&cfg, &e);
    if (rc) throw std::runtime_error(b200nuts_last_error(nullptr));
    std::vector<double> lp((size_t)n), g((size_t)n * D);
    if ((rc = model.upload(e)) || (rc = b200nuts_log_prob_grad(e, q.data(), n, jacobian ? 1 : 0, lp.data(), g.data()))) throw std::runtime_error(b200nuts_last_error(e));
    std::unique_ptr<std::ostream> out = create_stream(parser.str({"output", "file"}), sig_figs);
    char date[64];
    std::time_t now = std::time(nullptr);
    std::strftime(date, sizeof date, "%Y-%m-%d %H:%M:%S", std::gmtime(&now));
    *out << "# stan_version_major = 2\n# stan_version_minor = 36\n# stan_version_patch = 0\n# model = " << model.model_name() << "\n# start_datetime = " << date << " UTC\n";
    for (const std::string& ln : parser.echo()) *out << "# " << ln << "\n";
    for (const std::string& ln : model.model_compile_info()) *out << "# " << ln << "\n";
    std::vector<std::string> names;
    model.unconstrained_param_names(names, false, false);
    *out << "lp__";
    for (const std::string& nm : names) *out << ",g_" << nm;
    *out << "\n";
    for (int64_t i = 0; i < n; i++) {
      *out << lp[(size_t)i];
      for (int64_t k = 0; k < D; k++) *out << "," << g[(size_t)(i * D + k)];
      *out << "\n";
    }
    b200nuts_destroy(e);
    return 0;
  } catch (const std::domain_error& ex) {
    std::cerr << ex.what() << std::endl;
  } catch (const std::exception& ex) {
    std::cerr << ex.what() << std::endl;
  }
  if (e) b200nuts_destroy(e);
  return 1;
}

}  // namespace

int main(int argc, char** argv) {
  using namespace b200;
  std::vector<std::string> tokens;
  std::string model_name = model_from_argv0(argv[0]);
  int first = 1;
  if (model_name == "b200stan") {
    if (argc < 2) {
      std::cerr << "usage: b200stan <model> sample [arguments]   (models: std_normal bernoulli eight_schools_nc logistic irt_2pl poisson_glm,\n"
                   "        or a plug-in model of the libraries named in B200NUTS_MODEL_LIBRARY -- include/b200nuts_model.h)" << std::endl;
      return services::error_codes::USAGE;
    }
    model_name = argv[1];
    first = 2;
  }
  for (int i = first; i < argc; i++) tokens.push_back(argv[i]);
  callbacks::stream_logger logger(std::cout, std::cout, std::cout, std::cerr, std::cerr);
  cmd::argument_parser parser;
  if (tokens.empty() || tokens[0] == "help" || tokens[0] == "help-all") {  // argument_parser::parse_args help branches: exit 0
    std::cout << "Usage: " << argv[0] << (first == 2 ? std::string(" ") + argv[1] : std::string(""))
              << " <arg1> <subarg1_1> ... <subarg1_m> ... <arg_n> <subarg_n_1> ... <subarg_n_m>\n\n"
              << "Inference method available in the B200 engine:\n"
              << "  sample      Bayesian inference with Markov Chain Monte Carlo (algorithm=hmc engine=nuts metric=diag_e|unit_e|dense_e)\n\n"
              << "Or see help information with\n  help        Prints help\n  help-all    Prints entire argument tree\n\n"
              << "Additional configuration available by specifying\n  id, data, init, random, output, num_threads   (see help-all)\n";
    if (!tokens.empty() && tokens[0] == "help-all") {
      std::cout << "\n";
      for (const std::string& ln : parser.help_all()) std::cout << ln << "\n";
    }
    if (tokens.empty()) {
      std::cerr << "A method must be specified!\nFailed to parse arguments, terminating Stan" << std::endl;
      return services::error_codes::USAGE;
    }
    return 0;
  }
  try {
    parser.parse(tokens);
  } catch (const std::exception& e) {
    std::cerr << e.what() << std::endl << "Failed to parse arguments, terminating Stan" << std::endl;
    return services::error_codes::USAGE;
  }
  if (parser.str({"method"}) == "log_prob") return run_log_prob(parser, model_name);
  if (parser.str({"method"}) != "sample") {
    std::cerr << "method=" << parser.str({"method"}) << " is not available in the B200 engine (only `sample` with algorithm=hmc engine=nuts metric=diag_e|unit_e|dense_e)" << std::endl;
    return services::error_codes::USAGE;
  }
  const std::string metric = parser.str({"method", "sample", "algorithm", "hmc", "metric"});
  const std::string hmc_engine = parser.str({"method", "sample", "algorithm", "hmc", "engine"});
  if (parser.str({"method", "sample", "algorithm"}) != "hmc" || (hmc_engine != "nuts" && hmc_engine != "static") ||
      (metric != "diag_e" && metric != "unit_e" && metric != "dense_e")) {
    std::cerr << "only algorithm=hmc engine=nuts|static metric=diag_e|unit_e|dense_e is available in the B200 engine" << std::endl;
    return services::error_codes::USAGE;
  }
  const unsigned num_chains = (unsigned)parser.integer({"method", "sample", "num_chains"});
  const unsigned id = (unsigned)parser.integer({"id"});
  const int num_warmup = (int)parser.integer({"method", "sample", "num_warmup"});
  const int num_samples = (int)parser.integer({"method", "sample", "num_samples"});
  const int num_thin = (int)parser.integer({"method", "sample", "thin"});
  const bool save_warmup = parser.boolean({"method", "sample", "save_warmup"});
  const bool adapt_engaged = parser.boolean({"method", "sample", "adapt", "engaged"});
  const int sig_figs = (int)parser.integer({"output", "sig_figs"});
  const std::string output_file = parser.str({"output", "file"});
  if (adapt_engaged && num_warmup == 0) {  // command.hpp:569-574
    std::cerr << "The number of warmup samples (num_warmup) must be greater than zero if adaptation is enabled." << std::endl;
    return services::error_codes::USAGE;
  }
  if (parser.boolean({"method", "sample", "adapt", "save_metric"}) && !adapt_engaged) {
    std::cerr << "Saving the adapted metric (save_metric=1) can only be enabled when adaptation is engaged (adapt engaged=1)" << std::endl;
    return services::error_codes::USAGE;
  }
  // random seed: -1 => time based (arg_seed.hpp)
  long long seed_arg = parser.integer({"random", "seed"});
  unsigned int random_seed = seed_arg < 0 ? (unsigned int)(std::time(nullptr) % 4294967296ULL) : (unsigned int)seed_arg;
  if (seed_arg < 0) {
    parser.find({"random", "seed"})->value = std::to_string(random_seed);
    parser.find({"random", "seed"})->def = std::to_string(random_seed);  // echoed as "seed = <time-based value> (Default)", like the reference
  }

  // echo to the console (parser.print(info))
  for (const std::string& ln : parser.echo()) std::cout << ln << std::endl;
  std::cout << std::endl;

  int rc = services::error_codes::OK;
  try {
    io::var_context data = io::read_json_file(parser.str({"data", "file"}));
    model::device_model& model = model::new_model(model_name, data, random_seed, &std::cout);

    // inits: a number => radius, else a JSON file with constrained values (one file for all chains, or <base>_<id>.json)
    std::string init = parser.str({"init"});
    double init_radius = 2.0;
    {
      char* end = nullptr;
      double r = std::strtod(init.c_str(), &end);
      if (!init.empty() && *end == '\0') {
        init_radius = r;
        init = "";
      }
    }
    auto per_chain_contexts = [&](const std::string& file) {
      std::vector<std::shared_ptr<io::var_context>> v(num_chains);
      for (unsigned i = 0; i < num_chains; i++) {
        if (file.empty()) {
          v[i] = std::make_shared<io::var_context>();
          continue;
        }
        std::string path = file;
        if (num_chains > 1) {  // services_helper get_vec_var_context: prefer <base>_<id+i><suffix> when it exists
          auto bs = basename_suffix(file);
          std::string cand = bs.first + "_" + std::to_string(id + i) + bs.second;
          if (std::ifstream(cand).good()) path = cand;
        }
        v[i] = std::make_shared<io::var_context>(io::read_json_file(path));
      }
      return v;
    };
    auto init_contexts = per_chain_contexts(init);
    auto metric_contexts = per_chain_contexts(parser.str({"method", "sample", "algorithm", "hmc", "metric_file"}));

    // writers
    std::vector<callbacks::unique_stream_writer> sample_writers;
    std::vector<callbacks::writer> init_writers(num_chains);
    std::vector<callbacks::unique_stream_writer> diagnostic_writers;  // output diagnostic_file=: momenta and gradients of every saved draw
    const std::string diagnostic_file = parser.str({"output", "diagnostic_file"});
    if (!diagnostic_file.empty())
      for (const std::string& f : make_filenames(diagnostic_file, "", ".csv", num_chains, id)) diagnostic_writers.emplace_back(create_stream(f, sig_figs), "# ");
    std::vector<services::json_metric_writer> metric_writers;
    for (const std::string& f : make_filenames(output_file, "", ".csv", num_chains, id)) sample_writers.emplace_back(create_stream(f, sig_figs), "# ");
    if (parser.boolean({"method", "sample", "adapt", "save_metric"}))
      for (const std::string& f : make_filenames(output_file, "_metric", ".json", num_chains, id)) metric_writers.emplace_back(create_stream(f, sig_figs));

    // CSV header: write_config = write_stan + write_model + write_datetime + parser.print + compile info
    char date[64];
    std::time_t now = std::time(nullptr);
    std::strftime(date, sizeof date, "%Y-%m-%d %H:%M:%S", std::gmtime(&now));
    for (auto& w : sample_writers) {
      w("stan_version_major = 2");
      w("stan_version_minor = 36");
      w("stan_version_patch = 0");
      w("model = " + model.model_name());
      w(std::string("start_datetime = ") + date + " UTC");
      for (const std::string& ln : parser.echo()) w(ln);
      for (const std::string& ln : model.model_compile_info()) w(ln);
    }

    for (auto& w : diagnostic_writers) {  // command.hpp:288-291: write_stan + write_model + parser.print (no date, no compile info)
      w("stan_version_major = 2");
      w("stan_version_minor = 36");
      w("stan_version_patch = 0");
      w("model = " + model.model_name());
      for (const std::string& ln : parser.echo()) w(ln);
    }
    if (parser.boolean({"output", "save_cmdstan_config"})) {  // command.hpp:279-285 + write_config(structured_writer&)
      std::unique_ptr<std::ostream> cf = create_stream(basename_suffix(output_file).first + "_config.json", sig_figs);
      std::string js = "\n{\n  \"stan_major_version\" : \"2\",\n  \"stan_minor_version\" : \"36\",\n  \"stan_patch_version\" : \"0\",\n";
      js += "  \"model_name\" : \"" + model.model_name() + "\",\n  \"start_datetime\" : \"" + std::string(date) + " UTC\",\n";
      parser.echo_json(js, 1);
      js += "  \"mpi_enabled\" : false,\n";
      const std::vector<std::string> ci = model.model_compile_info();  // "stanc_version = ...", "stancflags = ..."
      for (size_t i = 0; i < ci.size(); i++) {
        const size_t eq = ci[i].find(" = ");
        const std::string key = ci[i].substr(0, eq), val = eq == std::string::npos ? "" : ci[i].substr(eq + 3);
        js += "  \"" + key + "\" : \"" + val + "\"" + (i + 1 < ci.size() ? ",\n" : "\n");
      }
      js += "}\n";
      *cf << js;
    }

    services::device_options opt;
    opt.unit_e = metric == "unit_e";  // command.hpp dispatches to hmc_nuts_unit_e[_adapt]; same call below with the identity metric
    opt.dense_e = metric == "dense_e";  // ... and to hmc_nuts_dense_e[_adapt] (command.hpp:620-660)
    opt.static_hmc = hmc_engine == "static";  // ... and to the six hmc_static_* services
    opt.int_time = parser.real({"method", "sample", "algorithm", "hmc", "engine", "static", "int_time"});
    if (opt.static_hmc && !diagnostic_writers.empty()) {
      std::cerr << "output diagnostic_file is available with engine=nuts only in the B200 engine" << std::endl;
      return services::error_codes::USAGE;
    }
    if (const char* dv = std::getenv("B200_DEVICES")) {
      opt.devices.clear();
      std::stringstream ss(dv);
      std::string t;
      while (std::getline(ss, t, ','))
        if (!t.empty()) opt.devices.push_back(std::atoi(t.c_str()));
      if (opt.devices.empty()) opt.devices.push_back(0);
    }
    callbacks::interrupt interrupt;
    rc = services::sample::hmc_nuts_diag_e_adapt(
        model, num_chains, init_contexts, metric_contexts, random_seed, id, init_radius, num_warmup, num_samples, num_thin, save_warmup,
        (int)parser.integer({"output", "refresh"}), parser.real({"method", "sample", "algorithm", "hmc", "stepsize"}),
        parser.real({"method", "sample", "algorithm", "hmc", "stepsize_jitter"}),
        (int)parser.integer({"method", "sample", "algorithm", "hmc", "engine", "nuts", "max_depth"}), parser.real({"method", "sample", "adapt", "delta"}),
        parser.real({"method", "sample", "adapt", "gamma"}), parser.real({"method", "sample", "adapt", "kappa"}), parser.real({"method", "sample", "adapt", "t0"}),
        (unsigned)parser.integer({"method", "sample", "adapt", "init_buffer"}), (unsigned)parser.integer({"method", "sample", "adapt", "term_buffer"}),
        (unsigned)parser.integer({"method", "sample", "adapt", "window"}), interrupt, logger, init_writers, sample_writers, diagnostic_writers,
        metric_writers, opt, adapt_engaged);
  } catch (const std::domain_error& e) {
    std::cerr << e.what() << std::endl;
    return services::error_codes::DATAERR;
  } catch (const std::exception& e) {
    std::cerr << e.what() << std::endl;
    return services::error_codes::USAGE;
  }
  return rc == 0 ? 0 : 1;  // main.cpp:7-10: any service error => process exit code 1
}
