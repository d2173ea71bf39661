// On-device posterior summary (summary.cu): the stansummary table computed from the draws an engine holds in HBM.
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "models.cuh"

namespace b200 {

// columns of the table, in stansummary's order (C/src/cmdstan/stansummary_helper.hpp:405-428)
constexpr int kSummaryStats = 10;  // Mean, MCSE, StdDev, MAD, 5%, 50%, 95%, ESS_bulk, ESS_tail, R_hat

struct SummaryArgs {
  const double* draws;   // [C][rows_per_chain][ncol] stored draws (sampler columns + unconstrained q)
  size_t draws_stride;   // doubles per chain
  int ncol;              // 7 + D
  int kind, D, N, K, K_out;
  int constrained;       // 1: columns 7.. are write_array outputs; 0: raw unconstrained q
  long long C, row0, n;  // chains, first row, rows per chain
  cudaStream_t stream;
};

// host_out [7 + K][kSummaryStats]; returns a stan::services::error_codes value
int summary_run(const SummaryArgs& a, double* host_out, std::string* err);

}  // namespace b200
