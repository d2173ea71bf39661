// CmdStan's command-line argument tree for `method=sample` (the caller of the path), restated:
//   grammar / defaults / validity   src/cmdstan/arguments/arg_sample.hpp:18-29, arg_adapt.hpp:19-37, arg_hmc.hpp:19-27,
//                                   arg_nuts.hpp:15-16, arg_engine.hpp:16-21, arg_metric.hpp:17-23, arg_id.hpp, arg_data.hpp,
//                                   arg_init.hpp, arg_random.hpp, arg_seed.hpp, arg_output.hpp:19-28, arg_num_threads.hpp
//   parsing                         src/cmdstan/arguments/argument_parser.hpp (sequential tokens; a token that is not a
//                                   sub-argument of the current category pops back to its ancestors)
//   echo                            argument::print: "name = value[ (Default)]", two spaces per depth, selected list
//                                   element printed as a nested category (SURVEY.md Appendix D golden header)
// Only `sample` is built here; the other methods are outside the hot path and are rejected by name.
#pragma once
#include <cmath>
#include <cstdlib>
#include <functional>
#include <limits>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace b200 {
namespace cmd {

struct arg {
  enum kind_t { CATEGORY, INT, UINT, LONG, REAL, BOOL, STRING, LIST } kind = CATEGORY;
  std::string name, value, def;
  std::string valid;  // description of the valid range for error messages
  std::function<bool(const arg&)> ok;
  std::vector<std::unique_ptr<arg>> kids;
  arg* parent = nullptr;
  double rval = 0.0;  // REAL: the parsed value at full precision (`value` holds the 6-digit echo)

  double def_rval = 0.0;  // REAL: the default at full precision (int_time defaults to 2 pi, echoed as 6.28319)
  bool is_default() const { return kind == REAL ? rval == def_rval : value == def; }
  arg* child(const std::string& n) {
    for (auto& k : kids)
      if (k->name == n) return k.get();
    return nullptr;
  }
  arg* add(std::unique_ptr<arg> a) {
    a->parent = this;
    kids.push_back(std::move(a));
    return kids.back().get();
  }
  double real() const { return kind == REAL ? rval : std::strtod(value.c_str(), nullptr); }
  long long integer() const { return std::strtoll(value.c_str(), nullptr, 10); }
  bool boolean() const { return value == "true" || value == "1"; }
};

inline std::unique_ptr<arg> mk(arg::kind_t k, const std::string& name, const std::string& def, const std::string& valid = "",
                               std::function<bool(const arg&)> ok = nullptr) {
  std::unique_ptr<arg> a(new arg());
  a->kind = k;
  a->name = name;
  a->value = a->def = def;
  a->rval = a->def_rval = std::strtod(def.c_str(), nullptr);
  a->valid = valid;
  a->ok = ok;
  return a;
}

class argument_parser {
 public:
  argument_parser() {
    root_.name = "";
    // method
    arg* method = root_.add(mk(arg::LIST, "method", "sample"));
    arg* sample = method->add(mk(arg::CATEGORY, "sample", ""));
    for (const char* other : {"optimize", "variational", "diagnose", "generate_quantities", "pathfinder"}) method->add(mk(arg::CATEGORY, other, ""));
    arg* lpm = method->add(mk(arg::CATEGORY, "log_prob", ""));  // arguments/arg_log_prob.hpp
    lpm->add(mk(arg::STRING, "unconstrained_params", ""));
    lpm->add(mk(arg::STRING, "constrained_params", ""));
    lpm->add(mk(arg::BOOL, "jacobian", "true"));
    method->add(mk(arg::CATEGORY, "laplace", ""));
    sample->add(mk(arg::INT, "num_samples", "1000", "0 <= num_samples", [](const arg& a) { return a.integer() >= 0; }));
    sample->add(mk(arg::INT, "num_warmup", "1000", "0 <= num_warmup", [](const arg& a) { return a.integer() >= 0; }));
    sample->add(mk(arg::BOOL, "save_warmup", "false"));
    sample->add(mk(arg::INT, "thin", "1", "0 < thin", [](const arg& a) { return a.integer() > 0; }));
    arg* adapt = sample->add(mk(arg::CATEGORY, "adapt", ""));
    adapt->add(mk(arg::BOOL, "engaged", "true"));
    adapt->add(mk(arg::REAL, "gamma", "0.05", "0 < gamma", [](const arg& a) { return a.real() > 0; }));
    adapt->add(mk(arg::REAL, "delta", "0.8", "0.1 <= delta <= 1", [](const arg& a) { return a.real() >= 0.1 && a.real() <= 1; }));
    adapt->add(mk(arg::REAL, "kappa", "0.75", "0 < kappa", [](const arg& a) { return a.real() > 0; }));
    adapt->add(mk(arg::REAL, "t0", "10", "0 < t0", [](const arg& a) { return a.real() > 0; }));
    adapt->add(mk(arg::UINT, "init_buffer", "75", "All", nullptr));
    adapt->add(mk(arg::UINT, "term_buffer", "50", "All", nullptr));
    adapt->add(mk(arg::UINT, "window", "25", "All", nullptr));
    adapt->add(mk(arg::BOOL, "save_metric", "false"));
    arg* algo = sample->add(mk(arg::LIST, "algorithm", "hmc"));
    arg* hmc = algo->add(mk(arg::CATEGORY, "hmc", ""));
    algo->add(mk(arg::CATEGORY, "fixed_param", ""));
    arg* engine = hmc->add(mk(arg::LIST, "engine", "nuts"));
    arg* stat = engine->add(mk(arg::CATEGORY, "static", ""));
    arg* int_time = stat->add(mk(arg::REAL, "int_time", "6.28319", "0 < int_time", [](const arg& a) { return a.real() > 0; }));
    int_time->rval = int_time->def_rval = 6.28318530717958647692;  // arguments/arg_int_time.hpp: 2 pi
    arg* nuts = engine->add(mk(arg::CATEGORY, "nuts", ""));
    nuts->add(mk(arg::INT, "max_depth", "10", "0 < max_depth", [](const arg& a) { return a.integer() > 0; }));
    arg* metric = hmc->add(mk(arg::LIST, "metric", "diag_e"));
    for (const char* m : {"unit_e", "diag_e", "dense_e"}) metric->add(mk(arg::CATEGORY, m, ""));
    hmc->add(mk(arg::STRING, "metric_file", ""));
    hmc->add(mk(arg::REAL, "stepsize", "1", "0 < stepsize", [](const arg& a) { return a.real() > 0; }));
    hmc->add(mk(arg::REAL, "stepsize_jitter", "0", "0 <= stepsize_jitter <= 1", [](const arg& a) { return a.real() >= 0 && a.real() <= 1; }));
    sample->add(mk(arg::INT, "num_chains", "1", "num_chains > 0", [](const arg& a) { return a.integer() > 0; }));
    // the rest of the top level
    root_.add(mk(arg::INT, "id", "1", "id >= 0", [](const arg& a) { return a.integer() >= 0; }));
    arg* data = root_.add(mk(arg::CATEGORY, "data", ""));
    data->add(mk(arg::STRING, "file", ""));
    root_.add(mk(arg::STRING, "init", "2"));
    arg* random = root_.add(mk(arg::CATEGORY, "random", ""));
    random->add(mk(arg::LONG, "seed", "-1", "non-negative integer < 4294967296  or -1",
                   [](const arg& a) { return a.integer() == -1 || (a.integer() >= 0 && a.integer() <= 4294967295LL); }));
    arg* output = root_.add(mk(arg::CATEGORY, "output", ""));
    output->add(mk(arg::STRING, "file", "output.csv"));
    output->add(mk(arg::STRING, "diagnostic_file", ""));
    output->add(mk(arg::INT, "refresh", "100"));
    output->add(mk(arg::INT, "sig_figs", "-1", "0 <= integer <= 18 or -1 to use the default number of significant figures",
                   [](const arg& a) { return a.integer() >= -1 && a.integer() <= 18; }));
    output->add(mk(arg::STRING, "profile_file", "profile.csv"));
    output->add(mk(arg::BOOL, "save_cmdstan_config", "false"));
    root_.add(mk(arg::INT, "num_threads", "1", "num_threads > 0 || num_threads == -1",
                 [](const arg& a) { return a.integer() > 0 || a.integer() == -1; }));
  }

  // returns 0 or throws std::invalid_argument with CmdStan's wording
  void parse(const std::vector<std::string>& tokens) {
    arg* ctx = &root_;
    bool method_seen = false;
    for (const std::string& tok : tokens) {
      std::string key = tok, val;
      const size_t eq = tok.find('=');
      const bool has_val = eq != std::string::npos;
      if (has_val) {
        key = tok.substr(0, eq);
        val = tok.substr(eq + 1);
      }
      arg* hit = nullptr;
      arg* scope = ctx;
      for (; scope; scope = scope->parent) {
        // a bare element name of a list argument in scope selects it ("sample", "hmc", "nuts")
        if ((hit = scope->child(key))) break;
        for (auto& k : scope->kids)
          if (k->kind == arg::LIST && !has_val && k->child(key)) {
            hit = k.get();
            val = key;
            break;
          }
        if (hit) break;
      }
      if (!hit) throw std::invalid_argument(tok + " is either mistyped or misplaced.");
      if (hit->kind == arg::CATEGORY && hit->parent && hit->parent->kind == arg::LIST && !has_val) {  // element of a list
        val = hit->name;
        hit = hit->parent;
      }
      if (hit->kind == arg::CATEGORY) {
        if (has_val) throw std::invalid_argument(tok + " is either mistyped or misplaced.");
        ctx = hit;
        continue;
      }
      if (hit->kind == arg::LIST) {
        if (hit->name == "method") method_seen = true;
        if (!has_val && val.empty()) throw std::invalid_argument("Value for " + key + " is missing");
        arg* sel = hit->child(val);
        if (!sel) throw std::invalid_argument(val + " is not a valid value for \"" + hit->name + "\"");
        hit->value = val;
        ctx = sel;
        continue;
      }
      if (!has_val) throw std::invalid_argument("Value for " + key + " is missing");
      set_value(hit, val);
      ctx = hit->parent;
    }
    if (!method_seen) throw std::invalid_argument("A method must be specified!");
  }

  arg* find(std::initializer_list<const char*> path) {
    arg* a = &root_;
    for (const char* p : path) {
      a = a->child(p);
      if (!a) throw std::logic_error(std::string("no such argument: ") + p);
    }
    return a;
  }
  std::string str(std::initializer_list<const char*> p) { return find(p)->value; }
  double real(std::initializer_list<const char*> p) { return find(p)->real(); }
  long long integer(std::initializer_list<const char*> p) { return find(p)->integer(); }
  bool boolean(std::initializer_list<const char*> p) { return find(p)->boolean(); }

  // config echo: lines without prefix (the caller adds "# " for CSV files)
  std::vector<std::string> echo() const {
    std::vector<std::string> out;
    for (auto& k : root_.kids) print(*k, 0, out);
    return out;
  }

  // `help-all`: the whole argument tree with types, defaults and valid values (the reference's argument_parser::help_printed
  // prints the same tree with its own prose; names, nesting, defaults and ranges are the contract)
  std::vector<std::string> help_all() const {
    std::vector<std::string> out;
    for (auto& k : root_.kids) help(*k, 0, out);
    return out;
  }

  // the same tree as members of a JSON object (argument_parser::print(structured_writer&), used by write_config for
  // `output save_cmdstan_config=1`, src/cmdstan/write_config.hpp:31-47): list arguments become {"value": v, v: {...}},
  // categories become objects, scalars JSON scalars.  Lines are appended to `out` at `depth` levels of two-space indentation,
  // each entry ending with a comma.
  void echo_json(std::string& out, int depth) const {
    for (auto& k : root_.kids) print_json(*k, depth, out);
  }

 private:
  static void set_value(arg* a, const std::string& v) {
    char* end = nullptr;
    std::string canon = v;
    switch (a->kind) {
      case arg::INT:
      case arg::UINT:
      case arg::LONG: {
        long long x = std::strtoll(v.c_str(), &end, 10);
        if (v.empty() || *end) throw std::invalid_argument("Invalid numeric value for " + a->name);
        if (a->kind == arg::UINT && x < 0) throw std::invalid_argument(v + " is not a valid value for \"" + a->name + "\"\n  Valid values: All");
        canon = std::to_string(x);
        break;
      }
      case arg::REAL: {
        double x = std::strtod(v.c_str(), &end);
        if (v.empty() || *end) throw std::invalid_argument("Invalid numeric value for " + a->name);
        std::stringstream ss;
        ss << x;  // printed the way the reference prints a double argument (operator<<, 6 significant digits)
        canon = ss.str();
        const double old = a->rval;
        a->rval = x;
        if (a->ok && !a->ok(*a)) {
          a->rval = old;
          throw std::invalid_argument(v + " is not a valid value for \"" + a->name + "\"\n  Valid values: " + a->valid);
        }
        a->value = canon;
        return;
      }
      case arg::BOOL:
        if (v == "1" || v == "true") canon = "true";
        else if (v == "0" || v == "false") canon = "false";
        else throw std::invalid_argument(v + " is not a valid value for \"" + a->name + "\"\n  Valid values: [0, 1, false, true]");
        break;
      default:
        break;
    }
    a->value = canon;
    if (a->ok && !a->ok(*a)) {
      a->value = a->def;
      throw std::invalid_argument(v + " is not a valid value for \"" + a->name + "\"\n  Valid values: " + a->valid);
    }
  }
  static void print(const arg& a, int depth, std::vector<std::string>& out) {
    const std::string ind(2 * depth, ' ');
    if (a.kind == arg::CATEGORY) {
      out.push_back(ind + a.name);
      for (auto& k : a.kids) print(*k, depth + 1, out);
      return;
    }
    out.push_back(ind + a.name + " = " + a.value + (a.is_default() ? " (Default)" : ""));
    if (a.kind == arg::LIST)
      for (auto& k : a.kids)
        if (k->name == a.value && !k->kids.empty()) print(*k, depth + 1, out);
  }
  static void help(const arg& a, int depth, std::vector<std::string>& out) {
    static const char* const kType[] = {"", "int", "unsigned int", "long long int", "double", "boolean", "string", "list element"};
    const std::string ind(2 * depth, ' ');
    if (a.kind == arg::CATEGORY) {
      out.push_back(ind + a.name);
    } else {
      out.push_back(ind + a.name + "=<" + kType[a.kind] + ">");
      if (a.kind == arg::LIST) {
        std::string v;
        for (auto& k : a.kids) v += (v.empty() ? "" : ", ") + k->name;
        out.push_back(ind + "  Valid values: " + v);
      } else if (!a.valid.empty()) {
        out.push_back(ind + "  Valid values: " + a.valid);
      }
      out.push_back(ind + "  Defaults to " + a.def);
    }
    for (auto& k : a.kids) help(*k, depth + 1, out);
  }
  static std::string json_string(const std::string& v) {
    std::string r = "\"";
    for (char ch : v) {
      if (ch == '"' || ch == '\\' || ch == '/') r += '\\';  // stan::callbacks::json_writer escapes the solidus too
      r += ch;
    }
    return r + "\"";
  }
  static void print_json(const arg& a, int depth, std::string& out) {
    const std::string ind(2 * depth, ' ');
    auto kids_of = [&](const arg& parent, int d) {
      for (auto& k : parent.kids) print_json(*k, d, out);
    };
    if (a.kind == arg::CATEGORY) {
      out += ind + "\"" + a.name + "\" : {\n";
      kids_of(a, depth + 1);
      close(out, ind);
      return;
    }
    if (a.kind == arg::LIST) {
      out += ind + "\"" + a.name + "\" : {\n" + ind + "  \"value\" : " + json_string(a.value) + ",\n";
      for (auto& k : a.kids)
        if (k->name == a.value && !k->kids.empty()) {
          out += ind + "  \"" + k->name + "\" : {\n";
          kids_of(*k, depth + 2);
          close(out, ind + "  ");
        }
      close(out, ind);
      return;
    }
    std::string v;
    if (a.kind == arg::BOOL) v = a.boolean() ? "true" : "false";
    else if (a.kind == arg::STRING) v = json_string(a.value);
    else v = a.value;
    out += ind + "\"" + a.name + "\" : " + v + ",\n";
  }
  // replaces the trailing ",\n" of the last member by "\n", then closes the object (itself a member: trailing comma)
  static void close(std::string& out, const std::string& ind) {
    if (out.size() >= 2 && out.compare(out.size() - 2, 2, ",\n") == 0) out.erase(out.size() - 2, 1);
    out += ind + "},\n";
  }
  arg root_;
};

}  // namespace cmd
}  // namespace b200
