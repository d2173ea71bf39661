// On-device posterior summary of the draws an engine holds in HBM: the table `stansummary` prints
// (Mean, MCSE, StdDev, MAD, 5%, 50%, 95%, ESS_bulk, ESS_tail, R_hat) for every output column, computed where the
// draws live so that thousands of chains never go through text CSV.
//
// Restates, for [chains x draws] data resident on the GPU:
//   split_chains                S/analyze/mcmc/split_chains.hpp:49-67
//   is_finite_and_varies        S/analyze/mcmc/check_chains.hpp:27-48
//   rank_transform              S/analyze/mcmc/rank_normalization.hpp:23-57   (average ranks, (r - 3/8) / (S + 1/4))
//   autocovariance (biased)     S/analyze/mcmc/autocovariance.hpp:32-88
//   ess (Geyer)                 S/analyze/mcmc/ess.hpp:28-103
//   split_rank_normalized_ess   S/analyze/mcmc/split_rank_normalized_ess.hpp:28-55
//   rhat / split_rank_..._rhat  S/analyze/mcmc/rhat.hpp:21-45, split_rank_normalized_rhat.hpp:27-43
//   mcse_mean                   S/analyze/mcmc/mcse.hpp:23-30
//   chainset mean/sd/mad/quantiles   S/mcmc/chainset.hpp:190-360, M/prim/fun/quantile.hpp:37-62
//   the table layout            C/src/cmdstan/stansummary_helper.hpp:405-428
//
// GPU formulation (not a translation): one column at a time; the pooled sort behind ranks and quantiles is a device
// radix sort (cub) of (value, position) pairs, tie groups are found by binary search in the sorted keys, the per-chain
// autocovariance is evaluated directly (sum of lagged products from a shared-memory copy of the chain, one CTA per
// chain) in chunks of lags -- Geyer's truncation usually stops inside the first chunk, so the O(n^2) tail is never
// computed -- and every reduction runs in a fixed order (deterministic table).
#include <cub/device/device_radix_sort.cuh>

#include <cmath>
#include <string>
#include <vector>

#include "summary.h"

namespace b200 {

namespace {

constexpr int kLagChunk = 128;
constexpr int kMaxChainLen = 24000;  // one chain of doubles in shared memory (192 KB)

struct ColScalars {
  // pooled statistics of the column (all draws of all chains)
  double mean, var, q05, q50, q95, mad;
  // split draws
  double s_q05, s_q50, s_q95;
  int ok_full, ok_split;
  // ess state machine (Geyer): persists across lag chunks
  double rho_even, rho_odd, w, var_plus;
  long long t;
  int done;
  double ess;
  // rhat of the last analysed matrix
  double rhat;
  // results
  double ess_bulk, ess_tail05, ess_tail95, ess_raw, rhat_bulk, rhat_tail;
};

// ---- gather: column k of the stored draws -> A [C][n] (all draws) and, when n is odd, S [2C][h] (split halves)
__global__ void gather_kernel(const double* __restrict__ draws, size_t draws_stride, int ncol, int kind, int D, int N, int K,
                              int constrained, int col, long long C, long long row0, long long n, double* __restrict__ A,
                              double* __restrict__ S) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= C * n) return;
  const long long c = idx / n, r = idx - c * n;
  const double* src = draws + (size_t)c * draws_stride + (size_t)(row0 + r) * ncol;
  double v;
  if (col < 7) v = src[col];
  else v = constrained ? model_output_column(kind, D, N, K, src + 7, col - 7) : src[col];
  A[idx] = v;
  if (S) {
    const long long h = n / 2, tail0 = (n + 1) / 2;
    if (r < h) S[(2 * c) * h + r] = v;
    else if (r >= tail0) S[(2 * c + 1) * h + (r - tail0)] = v;
  }
}

__global__ void iota_kernel(unsigned* idx, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) idx[i] = (unsigned)i;
}

// ---- block reduction in a fixed order (tree over shared memory)
template <int NT>
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int tid = threadIdx.x;
  red[tid] = v;
  __syncthreads();
#pragma unroll
  for (int s = NT / 2; s > 0; s >>= 1) {
    if (tid < s) red[tid] += red[tid + s];
    __syncthreads();
  }
  const double r = red[0];
  __syncthreads();
  return r;
}

// ---- is_finite_and_varies, per chain: flags[m] = 1 finite and not approx. constant, first[m] = first draw
__device__ __forceinline__ bool approx_equal(double a, double b) {  // Eigen internal::isApprox(a, b, 1e-12)
  return fabs(a - b) <= fmin(fabs(a), fabs(b)) * 1e-12;
}
__global__ void check_chain_kernel(const double* __restrict__ X, long long len, int* __restrict__ flags, double* __restrict__ first) {
  __shared__ int s_bad, s_varies;
  const long long m = blockIdx.x;
  const double* x = X + m * len;
  if (threadIdx.x == 0) { s_bad = 0; s_varies = 0; }
  __syncthreads();
  const double f = x[0];
  int bad = 0, varies = 0;
  for (long long i = threadIdx.x; i < len; i += blockDim.x) {
    const double v = x[i];
    if (!isfinite(v)) bad = 1;
    if (!approx_equal(v, f)) varies = 1;
  }
  if (bad) atomicOr(&s_bad, 1);
  if (varies) atomicOr(&s_varies, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    flags[m] = (!s_bad && s_varies) ? 1 : 0;
    first[m] = f;
  }
}
__global__ void check_combine_kernel(const int* __restrict__ flags, const double* __restrict__ first, long long M, int* out) {
  __shared__ int s_bad, s_varies;
  if (threadIdx.x == 0) { s_bad = 0; s_varies = 0; }
  __syncthreads();
  const double f0 = first[0];
  int bad = 0, varies = 0;
  for (long long m = threadIdx.x; m < M; m += blockDim.x) {
    if (!flags[m]) bad = 1;
    if (!approx_equal(first[m], f0)) varies = 1;
  }
  if (bad) atomicOr(&s_bad, 1);
  if (varies) atomicOr(&s_varies, 1);
  __syncthreads();
  if (threadIdx.x == 0) *out = (!s_bad && (M == 1 || s_varies)) ? 1 : 0;
}

// ---- pooled moments: deterministic two-stage sums
template <int MODE>  // 0: sum x ; 1: sum (x - c)^2
__global__ void partial_sum_kernel(const double* __restrict__ x, long long n, const double* centre, double* __restrict__ part) {
  __shared__ double red[256];
  const double c = MODE == 1 ? *centre : 0.0;
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const double v = x[i];
    s += MODE == 0 ? v : (v - c) * (v - c);
  }
  s = block_sum<256>(s, red);
  if (threadIdx.x == 0) part[blockIdx.x] = s;
}
__global__ void final_sum_kernel(const double* __restrict__ part, int nparts, double scale, double* out) {
  __shared__ double red[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < nparts; i += 256) s += part[i];
  s = block_sum<256>(s, red);
  if (threadIdx.x == 0) *out = s * scale;
}

// ---- quantiles of a sorted array (M/prim/fun/quantile.hpp:37-62)
__device__ __forceinline__ double sorted_quantile(const double* keys, long long n, double p) {
  if (n == 1) return keys[0];
  const double index = (double)(n - 1) * p;
  const long long lo = (long long)floor(index), hi = (long long)ceil(index);
  const double h = index - (double)lo;
  return (1 - h) * keys[lo] + h * keys[hi];
}
__global__ void quantile_kernel(const double* __restrict__ keys, long long n, double* q05, double* q50, double* q95) {
  if (q05) *q05 = sorted_quantile(keys, n, 0.05);
  if (q50) *q50 = sorted_quantile(keys, n, 0.5);
  if (q95) *q95 = sorted_quantile(keys, n, 0.95);
}
__global__ void mad_kernel(const double* __restrict__ keys, long long n, double* mad) { *mad = 1.4826 * sorted_quantile(keys, n, 0.5); }

__global__ void fold_kernel(const double* __restrict__ x, long long n, const double* centre, double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = fabs(x[i] - *centre);
}

// ---- rank_transform: sorted (key, position) pairs -> normal scores at the original positions
__global__ void rank_kernel(const double* __restrict__ keys, const unsigned* __restrict__ pos, long long n, double* __restrict__ Z) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double key = keys[i];
  long long lo = i, hi = i + 1;  // tie group [lo, hi)
  if (i > 0 && keys[i - 1] == key) {
    long long a = 0, b = i;  // first position whose key == key
    while (a < b) {
      const long long mid = (a + b) >> 1;
      if (keys[mid] < key) a = mid + 1;
      else b = mid;
    }
    lo = a;
  }
  if (i + 1 < n && keys[i + 1] == key) {
    long long a = i + 1, b = n;  // first position whose key > key
    while (a < b) {
      const long long mid = (a + b) >> 1;
      if (keys[mid] <= key) a = mid + 1;
      else b = mid;
    }
    hi = a;
  }
  const double avg_rank = ((double)(lo + 1) + (double)hi) / 2.0;  // mean of the ranks lo+1 .. hi
  const double p = (avg_rank - 0.375) / ((double)n + 0.25);
  Z[pos[i]] = normcdfinv(p);
}

// ---- per-chain autocovariance, lags [t0, t1): one CTA per chain, the chain centred in shared memory.
// MODE 0: x ; 1: indicator x <= thr ; 2: indicator x >= thr.
// A chain without variance gets NaN at every lag, as the reference's FFT route does (autocorrelation divides by ac(0)).
template <int MODE>
__global__ void __launch_bounds__(kLagChunk) acov_kernel(const double* __restrict__ X, int len, int t0, int t1, const double* thr,
                                                         double* __restrict__ ACOV, double* __restrict__ chain_mean,
                                                         double* __restrict__ chain_acov0) {
  extern __shared__ double xc[];
  __shared__ double red[kLagChunk];
  const long long m = blockIdx.x;
  const double* x = X + m * (long long)len;
  const int tid = threadIdx.x;
  const double th = MODE ? *thr : 0.0;
  double s = 0.0;
  for (int i = tid; i < len; i += kLagChunk) {
    double v = x[i];
    if (MODE == 1) v = v <= th ? 1.0 : 0.0;
    if (MODE == 2) v = v >= th ? 1.0 : 0.0;
    xc[i] = v;
    s += v;
  }
  const double mean = block_sum<kLagChunk>(s, red) / len;
  double s2 = 0.0;
  for (int i = tid; i < len; i += kLagChunk) {
    const double v = xc[i] - mean;
    xc[i] = v;
    s2 += v * v;
  }
  const double acov0 = block_sum<kLagChunk>(s2, red) / len;
  if (tid == 0 && t0 == 0) {
    chain_mean[m] = mean;
    chain_acov0[m] = acov0;
  }
  const int t = t0 + tid;
  if (t < t1) {
    double a = 0.0;
    const int cnt = len - t;
    for (int i = 0; i < cnt; i++) a += xc[i] * xc[i + t];
    ACOV[m * kLagChunk + tid] = acov0 == 0.0 ? NAN : a / len;
  }
}

// mean over chains of the chunk's autocovariances: one warp per lag, fixed order
__global__ void acov_mean_kernel(const double* __restrict__ ACOV, long long M, int t0, int t1, double* __restrict__ mean_acov) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int t = t0 + warp;
  if (t >= t1) return;
  double s = 0.0;
  for (long long m = lane; m < M; m += 32) s += ACOV[m * kLagChunk + warp];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) mean_acov[t] = s / (double)M;
}

// across-chain pieces of ess (W, var_plus) and rhat from the per-chain means / variances: one CTA, fixed order
__global__ void chain_combine_kernel(const double* __restrict__ chain_mean, const double* __restrict__ chain_acov0, long long M, int len,
                                     ColScalars* sc) {
  __shared__ double red[256];
  const int tid = threadIdx.x;
  double s = 0.0, sv = 0.0, sv_ess = 0.0;
  for (long long m = tid; m < M; m += 256) {
    s += chain_mean[m];
    const double cv = chain_acov0[m] * len / (len - 1.0);  // chain_var(i), ess.hpp:46
    sv += cv;
    sv_ess += chain_acov0[m] == 0.0 ? NAN : cv;  // the reference's FFT route yields NaN for a chain without variance
  }
  const double across = block_sum<256>(s, red) / (double)M;
  const double w_rhat = block_sum<256>(sv, red) / (double)M;
  const double w = block_sum<256>(sv_ess, red) / (double)M;
  double sb = 0.0;
  for (long long m = tid; m < M; m += 256) {
    const double d = chain_mean[m] - across;
    sb += d * d;
  }
  sb = block_sum<256>(sb, red);
  if (tid == 0) {
    const double var_means = M > 1 ? sb / (double)(M - 1) : 0.0;
    double var_plus = w * (len - 1.0) / len;
    if (M > 1) var_plus += var_means;
    sc->w = w;
    sc->var_plus = var_plus;
    // rhat.hpp:21-45
    const double between = (double)len * var_means;
    sc->rhat = sqrt((between / w_rhat + len - 1.0) / len);
    sc->t = 0;
    sc->done = 0;
  }
}

// Geyer's initial monotone sequence over the lags available so far (ess.hpp:60-101); resumable across chunks
__global__ void geyer_kernel(const double* __restrict__ mean_acov, double* __restrict__ rho, long long len, long long t_avail, long long M,
                             ColScalars* sc) {
  if (sc->done) return;
  const double w = sc->w, var_plus = sc->var_plus;
  long long t = sc->t;
  double rho_even = sc->rho_even, rho_odd = sc->rho_odd;
  if (t == 0) {
    rho_even = 1.0;
    rho[0] = rho_even;
    rho_odd = 1 - (w - mean_acov[1]) / var_plus;
    rho[1] = rho_odd;
    t = 1;
  }
  bool finished = true;
  while (t < len - 4 && (rho_even + rho_odd > 0) && !isnan(rho_even + rho_odd)) {
    if (t + 2 >= t_avail) {  // needs lags that are not computed yet
      finished = false;
      break;
    }
    rho_even = 1 - (w - mean_acov[t + 1]) / var_plus;
    rho_odd = 1 - (w - mean_acov[t + 2]) / var_plus;
    if (rho_even + rho_odd >= 0) {
      rho[t + 1] = rho_even;
      rho[t + 2] = rho_odd;
    }
    if (rho[t + 1] + rho[t + 2] > rho[t - 1] + rho[t]) {
      rho[t + 1] = (rho[t - 1] + rho[t]) / 2;
      rho[t + 2] = rho[t + 1];
    }
    t += 2;
  }
  sc->t = t;
  sc->rho_even = rho_even;
  sc->rho_odd = rho_odd;
  if (!finished) return;
  const long long max_t = t;
  if (rho_even > 0) rho[max_t + 1] = rho_even;
  const double total = (double)M * (double)len;
  double sum = 0.0;
  for (long long i = 0; i < max_t; i++) sum += rho[i];
  double tau = -1 + 2 * sum + rho[max_t + 1];
  tau = fmax(tau, 1 / log10(total));
  sc->ess = total / tau;
  sc->done = 1;
}

__global__ void store_kernel(ColScalars* sc, int what) {
  switch (what) {
    case 0: sc->ess_bulk = sc->ess; sc->rhat_bulk = sc->rhat; break;
    case 1: sc->ess_tail05 = sc->ess; break;
    case 2: sc->ess_tail95 = sc->ess; break;
    case 3: sc->ess_raw = sc->ess; break;
    case 4: sc->rhat_tail = sc->rhat; break;
  }
}

__global__ void finish_column_kernel(const ColScalars* sc, long long n_split, long long n_full, double* out) {
  const double nan = NAN;
  out[0] = sc->mean;
  out[1] = (n_full >= 4 && sc->ok_full) ? sqrt(sc->var / sc->ess_raw) : nan;  // mcse_mean
  out[2] = sqrt(sc->var);
  out[3] = sc->mad;
  out[4] = sc->q05;
  out[5] = sc->q50;
  out[6] = sc->q95;
  const bool ess_ok = sc->ok_split && n_split >= 4;
  out[7] = ess_ok ? sc->ess_bulk : nan;
  double tail = nan;
  if (ess_ok) {
    const double a = sc->ess_tail05, b = sc->ess_tail95;
    tail = isnan(a) ? b : (isnan(b) ? a : fmin(a, b));
  }
  out[8] = tail;
  out[9] = sc->ok_split ? fmax(sc->rhat_bulk, sc->rhat_tail) : nan;
}

struct Scratch {
  std::vector<void*> ptrs;
  std::string* err;
  template <typename T>
  bool alloc(T** p, size_t count) {
    void* q = nullptr;
    cudaError_t r = cudaMalloc(&q, count * sizeof(T) + 16);
    if (r != cudaSuccess) {
      *err = std::string("summary: cudaMalloc failed: ") + cudaGetErrorString(r);
      return false;
    }
    ptrs.push_back(q);
    *p = static_cast<T*>(q);
    return true;
  }
  ~Scratch() {
    for (void* p : ptrs) cudaFree(p);
  }
};

}  // namespace

int summary_num_stats() { return kSummaryStats; }

int summary_run(const SummaryArgs& a, double* host_out, std::string* err) {
  const long long C = a.C, n = a.n, h = n / 2, M = 2 * C;
  const long long n_full = C * n, n_split = M * h;
  if (n < 2) { *err = "summary: at least 2 draws per chain are needed"; return 64; }
  if (n > kMaxChainLen) { *err = "summary: more than 24000 draws per chain are not supported"; return 78; }
  if (n_full > 2147483647ll) { *err = "summary: more than 2^31 - 1 draws per column are not supported"; return 78; }  // cub's item counts are int
  const int ncols = 7 + (a.constrained ? a.K_out : a.D);
  const bool odd = (n & 1) != 0;
  cudaStream_t st = a.stream;

  Scratch sx;
  sx.err = err;
  double *A, *S = nullptr, *keys, *Z, *tmpv, *ACOV, *chain_mean, *chain_acov0, *first, *mean_acov, *rho, *part, *d_out;
  unsigned *pos_in, *pos_out;
  int* flags;
  ColScalars* sc;
  void* cub_tmp;
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const double*)nullptr, (double*)nullptr, (const unsigned*)nullptr, (unsigned*)nullptr, (int)n_full, 0, 64, st);
  const int nparts = 1024;
  if (!sx.alloc(&A, n_full) || (odd && !sx.alloc(&S, n_split)) || !sx.alloc(&keys, n_full) || !sx.alloc(&Z, n_full) || !sx.alloc(&tmpv, n_full) ||
      !sx.alloc(&ACOV, (size_t)std::max(M, C) * kLagChunk) || !sx.alloc(&chain_mean, M) || !sx.alloc(&chain_acov0, M) || !sx.alloc(&first, M) ||
      !sx.alloc(&mean_acov, n + 2) || !sx.alloc(&rho, n + 4) || !sx.alloc(&part, nparts) || !sx.alloc(&d_out, (size_t)ncols * kSummaryStats) ||
      !sx.alloc(&pos_in, n_full) || !sx.alloc(&pos_out, n_full) || !sx.alloc(&flags, M) || !sx.alloc(&sc, 1) ||
      !sx.alloc((char**)&cub_tmp, cub_bytes))
    return 70;
  if (!odd) S = A;  // [C][n] == [C][2][h] == [2C][h]: the split matrix is the same memory
  auto blocks = [](long long cnt, int nt) { return (unsigned)((cnt + nt - 1) / nt); };
  iota_kernel<<<blocks(n_full, 256), 256, 0, st>>>(pos_in, n_full);

  const size_t max_dyn = (size_t)kMaxChainLen * sizeof(double);
  cudaFuncSetAttribute(acov_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_dyn);
  cudaFuncSetAttribute(acov_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_dyn);
  cudaFuncSetAttribute(acov_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_dyn);

  int* h_done = nullptr;
  if (cudaMallocHost(&h_done, sizeof(int)) != cudaSuccess) { *err = "summary: cudaMallocHost failed"; return 70; }

  // ess of X [Mx][len] (MODE / thr as in acov_kernel); leaves ess and rhat in sc
  auto run_ess = [&](const double* X, long long Mx, long long len, int mode, const double* thr, bool rhat_only) -> bool {
    cudaMemsetAsync(rho, 0, sizeof(double) * (len + 4), st);
    for (long long t0 = 0; t0 < len; t0 += kLagChunk) {
      const long long t1 = rhat_only ? 1 : std::min(len, t0 + kLagChunk);
      const size_t dyn = (size_t)len * sizeof(double);
      if (mode == 0) acov_kernel<0><<<(unsigned)Mx, kLagChunk, dyn, st>>>(X, (int)len, (int)t0, (int)t1, thr, ACOV, chain_mean, chain_acov0);
      else if (mode == 1) acov_kernel<1><<<(unsigned)Mx, kLagChunk, dyn, st>>>(X, (int)len, (int)t0, (int)t1, thr, ACOV, chain_mean, chain_acov0);
      else acov_kernel<2><<<(unsigned)Mx, kLagChunk, dyn, st>>>(X, (int)len, (int)t0, (int)t1, thr, ACOV, chain_mean, chain_acov0);
      if (t0 == 0) chain_combine_kernel<<<1, 256, 0, st>>>(chain_mean, chain_acov0, Mx, (int)len, sc);
      if (rhat_only) return true;
      acov_mean_kernel<<<blocks((t1 - t0) * 32, 256), 256, 0, st>>>(ACOV, Mx, (int)t0, (int)t1, mean_acov);
      geyer_kernel<<<1, 1, 0, st>>>(mean_acov, rho, len, t1, Mx, sc);
      cudaMemcpyAsync(h_done, &sc->done, sizeof(int), cudaMemcpyDeviceToHost, st);
      if (cudaStreamSynchronize(st) != cudaSuccess) return false;
      if (*h_done) return true;
    }
    return true;
  };
  auto sort_pairs = [&](const double* in, long long cnt) {
    size_t bytes = cub_bytes;
    cub::DeviceRadixSort::SortPairs(cub_tmp, bytes, in, keys, pos_in, pos_out, (int)cnt, 0, 64, st);
  };

  bool ok = true;
  for (int col = 0; col < ncols && ok; col++) {
    gather_kernel<<<blocks(n_full, 256), 256, 0, st>>>(a.draws, a.draws_stride, a.ncol, a.kind, a.D, a.N, a.K, a.constrained, col, C, a.row0, n, A,
                                                        odd ? S : nullptr);
    // is_finite_and_varies of the unsplit and of the split matrix
    check_chain_kernel<<<(unsigned)C, 256, 0, st>>>(A, n, flags, first);
    check_combine_kernel<<<1, 256, 0, st>>>(flags, first, C, &sc->ok_full);
    if (h >= 1) {
      check_chain_kernel<<<(unsigned)M, 256, 0, st>>>(S, h, flags, first);
      check_combine_kernel<<<1, 256, 0, st>>>(flags, first, M, &sc->ok_split);
    } else {
      cudaMemsetAsync(&sc->ok_split, 0, sizeof(int), st);
    }
    // pooled mean / variance (chainset.hpp:190-210)
    partial_sum_kernel<0><<<nparts, 256, 0, st>>>(A, n_full, nullptr, part);
    final_sum_kernel<<<1, 256, 0, st>>>(part, nparts, 1.0 / (double)n_full, &sc->mean);
    partial_sum_kernel<1><<<nparts, 256, 0, st>>>(A, n_full, &sc->mean, part);
    final_sum_kernel<<<1, 256, 0, st>>>(part, nparts, 1.0 / (double)(n_full - 1), &sc->var);
    // pooled quantiles and MAD
    sort_pairs(A, n_full);
    quantile_kernel<<<1, 1, 0, st>>>(keys, n_full, &sc->q05, &sc->q50, &sc->q95);
    fold_kernel<<<blocks(n_full, 256), 256, 0, st>>>(A, n_full, &sc->q50, tmpv);
    {
      size_t bytes = cub_bytes;
      cub::DeviceRadixSort::SortKeys(cub_tmp, bytes, tmpv, Z, (int)n_full, 0, 64, st);
      mad_kernel<<<1, 1, 0, st>>>(Z, n_full, &sc->mad);
    }
    // mcse_mean: ess of the unsplit chains
    if (n >= 4) {
      ok = run_ess(A, C, n, 0, nullptr, false);
      store_kernel<<<1, 1, 0, st>>>(sc, 3);
    }
    if (h >= 1) {
      // bulk: rank-normalised split draws
      if (odd) sort_pairs(S, n_split);  // n even: the pooled sort above already is the sort of the split matrix
      quantile_kernel<<<1, 1, 0, st>>>(keys, n_split, &sc->s_q05, &sc->s_q50, &sc->s_q95);
      rank_kernel<<<blocks(n_split, 256), 256, 0, st>>>(keys, pos_out, n_split, Z);
      if (h >= 4) {
        ok = ok && run_ess(Z, M, h, 0, nullptr, false);
        store_kernel<<<1, 1, 0, st>>>(sc, 0);
        // tail: indicators of the 5 % / 95 % quantiles of the split draws
        ok = ok && run_ess(S, M, h, 1, &sc->s_q05, false);
        store_kernel<<<1, 1, 0, st>>>(sc, 1);
        ok = ok && run_ess(S, M, h, 2, &sc->s_q95, false);
        store_kernel<<<1, 1, 0, st>>>(sc, 2);
      } else {
        run_ess(Z, M, h, 0, nullptr, true);
        store_kernel<<<1, 1, 0, st>>>(sc, 0);
      }
      // tail R-hat: ranks of the draws folded at the median
      fold_kernel<<<blocks(n_split, 256), 256, 0, st>>>(S, n_split, &sc->s_q50, tmpv);
      sort_pairs(tmpv, n_split);
      rank_kernel<<<blocks(n_split, 256), 256, 0, st>>>(keys, pos_out, n_split, Z);
      run_ess(Z, M, h, 0, nullptr, true);
      store_kernel<<<1, 1, 0, st>>>(sc, 4);
    }
    finish_column_kernel<<<1, 1, 0, st>>>(sc, h, n, d_out + (size_t)col * kSummaryStats);
  }
  cudaMemcpyAsync(host_out, d_out, sizeof(double) * ncols * kSummaryStats, cudaMemcpyDeviceToHost, st);
  cudaError_t r = cudaStreamSynchronize(st);
  if (r == cudaSuccess) r = cudaGetLastError();
  cudaFreeHost(h_done);
  if (r != cudaSuccess || !ok) {
    *err = std::string("summary: ") + cudaGetErrorString(r);
    return 70;
  }
  return 0;
}

}  // namespace b200
