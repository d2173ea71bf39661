// Minimal JSON reader for Stan data / init files and the var_context it feeds
// (the subset of stan::json::json_data, stan/src/stan/io/json/json_data.hpp, that the path needs: top-level object of
// numbers and rectangular nested arrays of numbers; "NaN", "Infinity", "-Infinity" strings as in CmdStan JSON).
// Arrays are stored row-major (last index fastest) -- the engine's C ABI takes row-major data.
#pragma once
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace b200 {
namespace io {

struct var {
  std::vector<double> vals;
  std::vector<int64_t> dims;  // empty: scalar
  bool is_int = true;
};

class var_context {
 public:
  bool contains_r(const std::string& n) const { return vars_.count(n) != 0; }
  bool contains_i(const std::string& n) const { return vars_.count(n) != 0 && vars_.at(n).is_int; }
  const std::vector<double>& vals_r(const std::string& n) const { return get(n).vals; }
  std::vector<int> vals_i(const std::string& n) const {
    const var& v = get(n);
    if (!v.is_int) throw std::invalid_argument("variable " + n + " is not integer valued");
    std::vector<int> out(v.vals.size());
    for (size_t i = 0; i < out.size(); i++) out[i] = (int)v.vals[i];
    return out;
  }
  const std::vector<int64_t>& dims_r(const std::string& n) const { return get(n).dims; }
  std::vector<std::string> names() const {
    std::vector<std::string> r;
    for (auto& kv : vars_) r.push_back(kv.first);
    return r;
  }
  bool empty() const { return vars_.empty(); }
  void add(const std::string& n, var v) { vars_[n] = std::move(v); }

 private:
  const var& get(const std::string& n) const {
    auto it = vars_.find(n);
    if (it == vars_.end()) throw std::invalid_argument("variable does not exist; processing stage=data initialization; variable name=" + n);
    return it->second;
  }
  std::map<std::string, var> vars_;
};

class json_parser {
 public:
  explicit json_parser(const std::string& text) : s_(text), p_(0) {}
  var_context parse() {
    var_context ctx;
    ws();
    if (p_ >= s_.size()) return ctx;  // empty file == empty context
    expect('{');
    ws();
    if (peek() == '}') { p_++; return ctx; }
    for (;;) {
      ws();
      std::string key = str();
      ws();
      expect(':');
      var v;
      std::vector<int64_t> dims;
      skip_ = false;
      value(v, dims, 0);
      v.dims = dims;
      if (!skip_) ctx.add(key, std::move(v));  // a top-level string (e.g. "metric_type" : "dense_e" of a saved metric file) is not data
      ws();
      if (peek() == ',') { p_++; continue; }
      expect('}');
      break;
    }
    return ctx;
  }

 private:
  char peek() const { return p_ < s_.size() ? s_[p_] : '\0'; }
  void ws() { while (p_ < s_.size() && std::isspace((unsigned char)s_[p_])) p_++; }
  void expect(char c) {
    if (peek() != c) fail(std::string("expected '") + c + "'");
    p_++;
  }
  [[noreturn]] void fail(const std::string& m) const {
    throw std::invalid_argument("Error in JSON parsing at offset " + std::to_string(p_) + ": " + m);
  }
  std::string str() {
    expect('"');
    std::string r;
    while (p_ < s_.size() && s_[p_] != '"') {
      if (s_[p_] == '\\') p_++;
      r += s_[p_++];
    }
    expect('"');
    return r;
  }
  void value(var& v, std::vector<int64_t>& dims, size_t depth) {
    ws();
    if (peek() == '[') {
      p_++;
      int64_t n = 0;
      ws();
      if (peek() == ']') {
        p_++;
      } else {
        for (;;) {
          value(v, dims, depth + 1);
          n++;
          ws();
          if (peek() == ',') { p_++; continue; }
          expect(']');
          break;
        }
      }
      if (dims.size() <= depth) dims.resize(depth + 1, -1);
      if (dims[depth] == -1) dims[depth] = n;
      else if (dims[depth] != n) fail("non-rectangular array");
      return;
    }
    if (peek() == '"') {
      std::string t = str();
      if (t == "NaN") v.vals.push_back(std::numeric_limits<double>::quiet_NaN());
      else if (t == "Infinity" || t == "Inf") v.vals.push_back(std::numeric_limits<double>::infinity());
      else if (t == "-Infinity" || t == "-Inf") v.vals.push_back(-std::numeric_limits<double>::infinity());
      else if (depth == 0) skip_ = true;
      else fail("string values are not allowed: " + t);
      v.is_int = false;
      return;
    }
    size_t b = p_;
    while (p_ < s_.size() && (std::isdigit((unsigned char)s_[p_]) || std::strchr("+-.eE", s_[p_]))) p_++;
    if (b == p_) fail("number expected");
    std::string tok = s_.substr(b, p_ - b);
    if (tok.find_first_of(".eE") != std::string::npos) v.is_int = false;
    v.vals.push_back(std::strtod(tok.c_str(), nullptr));
  }
  std::string s_;
  size_t p_;
  bool skip_ = false;
};

inline var_context read_json_file(const std::string& path) {
  if (path.empty()) return var_context();
  std::ifstream in(path);
  if (!in.good()) throw std::invalid_argument("Cannot open specified file, \"" + path + "\"");
  std::stringstream ss;
  ss << in.rdbuf();
  return json_parser(ss.str()).parse();
}

}  // namespace io
}  // namespace b200
