// Example plug-in models for libb200nuts (include/b200nuts_model.h): what a user writes in place
// of a stanc-generated model class.  Build (done by cmdstan_b200/_build.py build_example_plugins):
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --fmad=false -Xcompiler -fPIC -shared \
//        examples/plugin_models/example_models.cu -o examples/plugin_models/libb200models_example.so
//
// Two models:
//
//  "linreg"  -- a model the engine does NOT have built in:
//        data { int N; int K; matrix[N,K] X; vector[N] y; }
//        parameters { real alpha; vector[K] beta; real<lower=0> sigma; }
//        model { alpha ~ normal(0, 5); beta ~ normal(0, 2); sigma ~ normal(0, 1); y ~ normal(alpha + X * beta, sigma); }
//     unconstrained q = (alpha, beta[K], log sigma).  With propto = true, jacobian = true (what the sampler sees;
//     normal_lpdf: M/prim/prob/normal_lpdf.hpp:44-105, lb_constrain: M/rev/constraint/lb_constrain.hpp:97):
//        lp = -alpha^2/50 - sum beta^2/8 - sigma^2/2 + log sigma - N log sigma - sum r_i^2 / (2 sigma^2),  r = y - alpha - X beta
//     One CTA per requesting chain; thread t owns rows t, t + 128, ...; every partial is reduced in a fixed order.
//
//  "eight_schools_plugin" -- the SAME arithmetic as the built-in eight-schools model (it calls the engine's own device
//     function), so that the plug-in path can be checked bit for bit against the built-in lock-step path.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../cmdstan_b200/csrc/models.cuh"  // only for the eight-schools twin; "linreg" needs nothing from the engine
#include "../../include/b200nuts.h"
#include "../../include/b200nuts_model.h"

namespace {

const b200nuts_data_view* find(const b200nuts_data_view* d, int n, const char* name) {
  for (int i = 0; i < n; i++)
    if (std::strcmp(d[i].name, name) == 0) return d + i;
  return nullptr;
}
int fail(char* err, int errlen, int code, const std::string& msg) {
  std::snprintf(err, (size_t)errlen, "%s", msg.c_str());
  return code;
}

// ---------------------------------------------------------------------------------------------------------------------
// linreg
// ---------------------------------------------------------------------------------------------------------------------
struct LinReg {
  int N, K, device;
  double *X, *y;  // device: X [N][K] row-major, y [N]
};

constexpr int kLinThreads = 128;
constexpr int kLinMaxD = 34;  // K <= 32

__global__ void __launch_bounds__(kLinThreads) linreg_grad_kernel(LinReg m, const double* Q, double* G, double* zV, const int* list,
                                                                 const int* n_list, int n_fixed, int ld, int jacobian) {
  const int n = n_list ? *n_list : n_fixed;
  const int s = blockIdx.x;
  if (s >= n) return;
  const int c = list ? list[s] : s;
  const int K = m.K, D = K + 2;
  __shared__ double qs[kLinMaxD];
  __shared__ double part[kLinThreads][kLinMaxD];  // per thread: [0] sum r^2, [1] sum r, [2..] sum r x_k
  const int t = threadIdx.x;
  if (t < D) qs[t] = Q[(size_t)c * ld + t];
  __syncthreads();
  double acc[kLinMaxD];
  for (int k = 0; k < D; k++) acc[k] = 0.0;
  for (int i = t; i < m.N; i += kLinThreads) {
    const double* x = m.X + (size_t)i * K;
    double mu = qs[0];
    for (int k = 0; k < K; k++) mu += x[k] * qs[1 + k];
    const double r = m.y[i] - mu;
    acc[0] += r * r;
    acc[1] += r;
    for (int k = 0; k < K; k++) acc[2 + k] += r * x[k];
  }
  for (int k = 0; k < D; k++) part[t][k] = acc[k];
  __syncthreads();
  // fixed-order tree over the 128 threads, one statistic per thread column
  for (int stride = kLinThreads / 2; stride > 0; stride >>= 1) {
    for (int e = t; e < stride * D; e += kLinThreads) {
      const int row = e / D, k = e - row * D;
      part[row][k] += part[row + stride][k];
    }
    __syncthreads();
  }
  if (t == 0) {
    const double alpha = qs[0], ls = qs[K + 1], sigma = exp(ls);
    const double inv_s2 = 1.0 / (sigma * sigma);
    double sb = 0;
    for (int k = 0; k < K; k++) sb += qs[1 + k] * qs[1 + k];
    double lp = -alpha * alpha / 50.0 - sb / 8.0 - 0.5 * sigma * sigma - m.N * ls - 0.5 * part[0][0] * inv_s2;
    if (jacobian) lp += ls;
    double* g = G + (size_t)c * ld;
    g[0] = -(-alpha / 25.0 + part[0][1] * inv_s2);
    for (int k = 0; k < K; k++) g[1 + k] = -(-qs[1 + k] / 4.0 + part[0][2 + k] * inv_s2);
    g[K + 1] = -(-sigma * sigma - m.N + part[0][0] * inv_s2 + (jacobian ? 1.0 : 0.0));
    zV[c] = -lp;
  }
}

__global__ void linreg_write_array_kernel(int D, const double* q, long long n, long long ldq, double* out, long long ldo) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  for (int i = 0; i < D - 1; i++) out[s * ldo + i] = q[s * ldq + i];
  out[s * ldo + D - 1] = exp(q[s * ldq + D - 1]);
}

int linreg_create(const b200nuts_data_view* data, int nd, int device, void** model, int64_t* D, int64_t* K_out, char* err, int errlen) {
  const b200nuts_data_view* X = find(data, nd, "X");
  const b200nuts_data_view* y = find(data, nd, "y");
  if (!X || !y) return fail(err, errlen, B200NUTS_NOINPUT, "variable does not exist; processing stage=data initialization; variable name=" + std::string(X ? "y" : "X"));
  if (X->dtype != B200NUTS_F64 || X->nd != 2 || y->dtype != B200NUTS_F64 || y->nd != 1 || y->dims[0] != X->dims[0])
    return fail(err, errlen, B200NUTS_DATAERR, "X must be matrix[N,K] and y vector[N] of doubles");
  if (X->dims[1] < 1 || X->dims[1] > kLinMaxD - 2) return fail(err, errlen, B200NUTS_CONFIG, "1 <= K <= 32 in this example");
  auto* m = new LinReg{(int)X->dims[0], (int)X->dims[1], device, nullptr, nullptr};
  if (cudaSetDevice(device) != cudaSuccess || cudaMalloc(&m->X, sizeof(double) * (size_t)m->N * m->K + 8) != cudaSuccess ||
      cudaMalloc(&m->y, sizeof(double) * (size_t)m->N + 8) != cudaSuccess ||
      cudaMemcpy(m->X, X->host, sizeof(double) * (size_t)m->N * m->K, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(m->y, y->host, sizeof(double) * (size_t)m->N, cudaMemcpyHostToDevice) != cudaSuccess) {
    const std::string msg = std::string("CUDA: ") + cudaGetErrorString(cudaGetLastError());
    cudaFree(m->X);
    cudaFree(m->y);
    delete m;
    return fail(err, errlen, B200NUTS_SOFTWARE, msg);
  }
  *model = m;
  *D = m->K + 2;
  *K_out = m->K + 2;
  return B200NUTS_OK;
}
void linreg_destroy(void* model) {
  auto* m = (LinReg*)model;
  cudaSetDevice(m->device);
  cudaFree(m->X);
  cudaFree(m->y);
  delete m;
}
int linreg_log_prob_grad(void* model, const double* Q, double* G, double* zV, const int* list, const int* n_list, int n_fixed, int ld,
                         int jacobian, void* stream) {
  auto* m = (LinReg*)model;
  if (n_fixed <= 0) return 0;
  linreg_grad_kernel<<<n_fixed, kLinThreads, 0, (cudaStream_t)stream>>>(*m, Q, G, zV, list, n_list, n_fixed, ld, jacobian);
  return 1;
}
int linreg_write_array(void* model, const double* q, int64_t n, int64_t ldq, double* out, int64_t ldo, void* stream) {
  auto* m = (LinReg*)model;
  if (n <= 0) return 0;
  linreg_write_array_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(m->K + 2, q, n, ldq, out, ldo);
  return 1;
}
int linreg_output_name(void* model, int64_t k, char* buf, int buflen) {
  auto* m = (LinReg*)model;
  if (k == 0) std::snprintf(buf, (size_t)buflen, "alpha");
  else if (k <= m->K) std::snprintf(buf, (size_t)buflen, "beta.%lld", (long long)k);
  else std::snprintf(buf, (size_t)buflen, "sigma");
  return B200NUTS_OK;
}
// transform_inits (model_base.hpp:348): outputs are (alpha, beta[K], sigma); sigma has a lower bound 0 (lb_free = log)
int linreg_transform_inits(void* model, const double* x, double* q, char* err, int errlen) {
  auto* m = (LinReg*)model;
  const int K = m->K;
  for (int i = 0; i <= K; i++) q[i] = x[i] == x[i] ? x[i] : b200nuts_init_not_given(0);
  const double sigma = x[K + 1];
  if (sigma != sigma) {
    q[K + 1] = b200nuts_init_not_given(1);
  } else if (!(sigma >= 0)) {
    std::snprintf(err, (size_t)errlen, "lb_free: Bounded variable is %g, but must be greater than or equal to 0", sigma);
    return B200NUTS_CONFIG;
  } else {
    q[K + 1] = std::log(sigma);
  }
  return 0;
}

const b200nuts_model_vtable kLinReg = {B200NUTS_MODEL_ABI, "linreg", linreg_create, linreg_destroy, linreg_log_prob_grad, linreg_write_array,
                                       linreg_output_name, linreg_transform_inits};

// ---------------------------------------------------------------------------------------------------------------------
// eight_schools_plugin: the built-in density behind the plug-in interface (bit-for-bit check of the mechanism)
// ---------------------------------------------------------------------------------------------------------------------
struct Schools {
  b200::ModelData md;
  int device;
  double *y, *sigma;
};

__global__ void __launch_bounds__(128) schools_grad_kernel(b200::ModelData md, const double* Q, double* G, double* zV, const int* list,
                                                          const int* n_list, int n_fixed, int ld, int jacobian) {
  const int n = n_list ? *n_list : n_fixed;
  const int s = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (s >= n) return;
  const int c = list ? list[s] : s;
  const int lane = threadIdx.x & 31;
  const double* q = Q + (size_t)c * ld;
  double* g = G + (size_t)c * ld;
  double lp = b200::lp_grad_eight_schools(md, q, g, lane);
  __syncwarp();
  for (int i = lane; i < md.D; i += 32) g[i] = -g[i];
  __syncwarp();
  if (lane == 0) {
    if (!jacobian) {  // lb_constrain's log-Jacobian term: lp += log tau = q[1]
      lp -= q[1];
      g[1] += 1.0;
    }
    zV[c] = -lp;
  }
}
__global__ void schools_write_array_kernel(int D, int J, const double* q, long long n, long long ldq, double* out, long long ldo) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  b200::model_write_array(b200::MODEL_EIGHT_SCHOOLS, D, J, 0, q + s * ldq, out + s * ldo);
}
int schools_create(const b200nuts_data_view* data, int nd, int device, void** model, int64_t* D, int64_t* K_out, char* err, int errlen) {
  const b200nuts_data_view* y = find(data, nd, "y");
  const b200nuts_data_view* sg = find(data, nd, "sigma");
  if (!y || !sg) return fail(err, errlen, B200NUTS_NOINPUT, "y and sigma must be given");
  if (y->dtype != B200NUTS_F64 || sg->dtype != B200NUTS_F64 || y->nd != 1 || sg->nd != 1 || y->dims[0] != sg->dims[0])
    return fail(err, errlen, B200NUTS_DATAERR, "y and sigma must be vectors of doubles of equal size");
  const int J = (int)y->dims[0];
  for (int j = 0; j < J; j++)
    if (!(((const double*)sg->host)[j] > 0)) return fail(err, errlen, B200NUTS_DATAERR, "sigma must be positive");
  auto* m = new Schools();
  std::memset(&m->md, 0, sizeof m->md);
  m->device = device;
  cudaSetDevice(device);
  cudaMalloc(&m->y, 8 * J);
  cudaMalloc(&m->sigma, 8 * J);
  cudaMemcpy(m->y, y->host, 8 * J, cudaMemcpyHostToDevice);
  cudaMemcpy(m->sigma, sg->host, 8 * J, cudaMemcpyHostToDevice);
  if (cudaGetLastError() != cudaSuccess) {
    delete m;
    return fail(err, errlen, B200NUTS_SOFTWARE, "CUDA allocation failed");
  }
  m->md.kind = b200::MODEL_EIGHT_SCHOOLS;
  m->md.N = J;
  m->md.D = J + 2;
  m->md.yr = m->y;
  m->md.sigma = m->sigma;
  *model = m;
  *D = J + 2;
  *K_out = 2 * J + 2;
  return B200NUTS_OK;
}
void schools_destroy(void* model) {
  auto* m = (Schools*)model;
  cudaSetDevice(m->device);
  cudaFree(m->y);
  cudaFree(m->sigma);
  delete m;
}
int schools_log_prob_grad(void* model, const double* Q, double* G, double* zV, const int* list, const int* n_list, int n_fixed, int ld,
                          int jacobian, void* stream) {
  auto* m = (Schools*)model;
  if (n_fixed <= 0) return 0;
  schools_grad_kernel<<<(n_fixed + 3) / 4, 128, 0, (cudaStream_t)stream>>>(m->md, Q, G, zV, list, n_list, n_fixed, ld, jacobian);
  return 1;
}
int schools_write_array(void* model, const double* q, int64_t n, int64_t ldq, double* out, int64_t ldo, void* stream) {
  auto* m = (Schools*)model;
  if (n <= 0) return 0;
  schools_write_array_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(m->md.D, m->md.N, q, n, ldq, out, ldo);
  return 1;
}
const b200nuts_model_vtable kSchools = {B200NUTS_MODEL_ABI, "eight_schools_plugin", schools_create, schools_destroy, schools_log_prob_grad,
                                        schools_write_array, nullptr};

}  // namespace

extern "C" const b200nuts_model_vtable* b200nuts_model_entry(int index) {
  if (index == 0) return &kLinReg;
  if (index == 1) return &kSchools;
  return nullptr;
}
