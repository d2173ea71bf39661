// TEST-ONLY (compile check, tests/test_abi_and_host.py): instantiates every service mirror of cmdstan_b200/host/service.hpp with
// the writer types the front end uses, so that the templates nobody calls inside the repository still have to compile.
#include <memory>
#include <vector>

#include "../../cmdstan_b200/host/service.hpp"

using namespace b200;

int instantiate_all(model::device_model& model) {
  std::vector<std::shared_ptr<io::var_context>> init, metric;
  std::vector<callbacks::writer> init_w(1);
  std::vector<callbacks::unique_stream_writer> sample_w, diag_w;
  std::vector<services::json_metric_writer> metric_w;
  callbacks::interrupt interrupt;
  callbacks::stream_logger logger(std::cout, std::cout, std::cout, std::cerr, std::cerr);
  int rc = 0;
  namespace S = services::sample;
  rc += S::hmc_nuts_diag_e_adapt(model, 1, init, metric, 1u, 1u, 2.0, 10, 10, 1, false, 100, 1.0, 0.0, 10, 0.8, 0.05, 0.75, 10.0, 75u, 50u, 25u, interrupt, logger,
                                 init_w, sample_w, diag_w, metric_w);
  rc += S::hmc_nuts_diag_e(model, 1, init, metric, 1u, 1u, 2.0, 10, 10, 1, false, 100, 1.0, 0.0, 10, interrupt, logger, init_w, sample_w, diag_w);
  rc += S::hmc_nuts_dense_e_adapt(model, 1, init, metric, 1u, 1u, 2.0, 10, 10, 1, false, 100, 1.0, 0.0, 10, 0.8, 0.05, 0.75, 10.0, 75u, 50u, 25u, interrupt, logger,
                                  init_w, sample_w, diag_w, metric_w);
  rc += S::hmc_nuts_dense_e(model, 1, init, metric, 1u, 1u, 2.0, 10, 10, 1, false, 100, 1.0, 0.0, 10, interrupt, logger, init_w, sample_w, diag_w);
  rc += S::hmc_nuts_unit_e_adapt(model, 1, init, 1u, 1u, 2.0, 10, 10, 1, false, 100, 1.0, 0.0, 10, 0.8, 0.05, 0.75, 10.0, interrupt, logger, init_w, sample_w, diag_w,
                                 metric_w);
  rc += S::hmc_nuts_unit_e(model, 1, init, 1u, 1u, 2.0, 10, 10, 1, false, 100, 1.0, 0.0, 10, interrupt, logger, init_w, sample_w, diag_w);
  rc += S::hmc_static_diag_e_adapt(model, 1, init, metric, 1u, 1u, 2.0, 10, 10, 1, false, 100, 1.0, 0.0, 6.28, 0.8, 0.05, 0.75, 10.0, 75u, 50u, 25u, interrupt, logger,
                                   init_w, sample_w, diag_w);
  rc += S::hmc_static_diag_e(model, 1, init, metric, 1u, 1u, 2.0, 10, 10, 1, false, 100, 1.0, 0.0, 6.28, interrupt, logger, init_w,
                             sample_w, diag_w);
  rc += S::hmc_static_unit_e_adapt(model, 1, init, metric, 1u, 1u, 2.0, 10, 10, 1, false, 100, 1.0, 0.0, 6.28, 0.8, 0.05, 0.75, 10.0, 75u, 50u, 25u, interrupt, logger,
                                   init_w, sample_w, diag_w);
  rc += S::hmc_static_unit_e(model, 1, init, metric, 1u, 1u, 2.0, 10, 10, 1, false, 100, 1.0, 0.0, 6.28, interrupt, logger, init_w,
                             sample_w, diag_w);
  rc += S::hmc_static_dense_e_adapt(model, 1, init, metric, 1u, 1u, 2.0, 10, 10, 1, false, 100, 1.0, 0.0, 6.28, 0.8, 0.05, 0.75, 10.0, 75u, 50u, 25u, interrupt, logger,
                                    init_w, sample_w, diag_w);
  rc += S::hmc_static_dense_e(model, 1, init, metric, 1u, 1u, 2.0, 10, 10, 1, false, 100, 1.0, 0.0, 6.28, interrupt, logger, init_w,
                              sample_w, diag_w);
  return rc;
}
