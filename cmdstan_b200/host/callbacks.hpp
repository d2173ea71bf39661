// Host-side callback interfaces of the B200 NUTS service -- same shapes as the reference's
//   stan::callbacks::writer            stan/src/stan/callbacks/writer.hpp:16-62
//   stan::callbacks::unique_stream_writer   stan/src/stan/callbacks/unique_stream_writer.hpp:60-161
//   stan::callbacks::logger / stream_logger stan/src/stan/callbacks/logger.hpp, stream_logger.hpp
//   stan::callbacks::interrupt         stan/src/stan/callbacks/interrupt.hpp
// so that a caller written against the reference's service (cmdstan::command) ports by changing a namespace.
#pragma once
#include <iostream>
#include <memory>
#include <ostream>
#include <sstream>
#include <string>
#include <vector>

namespace b200 {
namespace callbacks {

class writer {
 public:
  virtual ~writer() {}
  virtual void operator()(const std::vector<std::string>&) {}
  virtual void operator()(const std::vector<double>&) {}
  virtual void operator()() {}
  virtual void operator()(const std::string&) {}
  virtual bool is_null() const { return true; }
  // access for bulk row output (rows are formatted by the service straight into the stream)
  virtual std::ostream* stream() { return nullptr; }
};

// owns its stream; comment lines get `comment_prefix`; values are comma separated at the stream's precision
class unique_stream_writer : public writer {
 public:
  explicit unique_stream_writer(std::unique_ptr<std::ostream>&& output, const std::string& comment_prefix = "")
      : output_(std::move(output)), comment_prefix_(comment_prefix) {}
  void operator()(const std::vector<std::string>& names) override { write_vector(names); }
  void operator()(const std::vector<double>& values) override { write_vector(values); }
  void operator()() override {
    if (output_) *output_ << comment_prefix_ << std::endl;
  }
  void operator()(const std::string& message) override {
    if (output_) *output_ << comment_prefix_ << message << std::endl;
  }
  bool is_null() const override { return output_ == nullptr; }
  std::ostream* stream() override { return output_.get(); }

 private:
  template <class T>
  void write_vector(const std::vector<T>& v) {
    if (!output_ || v.empty()) return;
    for (size_t i = 0; i + 1 < v.size(); ++i) *output_ << v[i] << ",";
    *output_ << v.back() << std::endl;
  }
  std::unique_ptr<std::ostream> output_;
  std::string comment_prefix_;
};

class logger {
 public:
  virtual ~logger() {}
  virtual void debug(const std::string&) {}
  virtual void info(const std::string&) {}
  virtual void warn(const std::string&) {}
  virtual void error(const std::string&) {}
  virtual void fatal(const std::string&) {}
  void info(const std::stringstream& s) { info(s.str()); }
  void error(const std::stringstream& s) { error(s.str()); }
};

class stream_logger : public logger {
 public:
  stream_logger(std::ostream& debug, std::ostream& info, std::ostream& warn, std::ostream& error, std::ostream& fatal)
      : debug_(debug), info_(info), warn_(warn), error_(error), fatal_(fatal) {}
  void debug(const std::string& m) override { debug_ << m << std::endl; }
  void info(const std::string& m) override { info_ << m << std::endl; }
  void warn(const std::string& m) override { warn_ << m << std::endl; }
  void error(const std::string& m) override { error_ << m << std::endl; }
  void fatal(const std::string& m) override { fatal_ << m << std::endl; }

 private:
  std::ostream &debug_, &info_, &warn_, &error_, &fatal_;
};

class interrupt {
 public:
  virtual ~interrupt() {}
  virtual void operator()() {}
};

}  // namespace callbacks
}  // namespace b200
