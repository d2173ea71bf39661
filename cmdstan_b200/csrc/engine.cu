// libb200nuts.so -- engine + C ABI (include/b200nuts.h).
//
// Two ways of running the per-chain state machine of nuts.cuh:
//
//  * fused engine (models 0-2: a gradient is tens of flops): one persistent warp per chain; the model's log density and
//    gradient are inlined, a chain runs all requested transitions without leaving the kernel.  One launch per
//    b200nuts_run.  Chain state lives in global memory but is private to its warp, so it stays in that SM's L1/L2.
//  * lock-step engine (GLMs, IRT: a gradient is 10^7..10^9 flops and shares the data matrix across chains): every chain
//    is advanced to its next "gradient wanted" point (advance_kernel, warp per chain), then ONE cross-chain kernel
//    evaluates all chains' gradients (glm.cu / irt.cu), and so on -- "asynchronous chains, synchronous leapfrogs".
//    Chains never wait for each other at transition boundaries; the only global synchronisation is the end of the run.
//
// Chain c of an engine uses the MIXMAX stream create_rng(seed, init_chain_id + c) and its output depends on nothing
// else (parallel-match invariant, ST/unit/services/sample/hmc_nuts_diag_e_adapt_parallel_match_test.cpp:73-153), which
// is what makes chain sharding across GPUs communication-free.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#define B200_WITH_DENSE 0  // dense-metric chains run the kernels of dense.cu
#include "../../include/b200nuts.h"
#include "../../include/b200nuts_model.h"
#include "host_seed.h"
#include "kernels.h"
#include "chains.cuh"
#include "nuts.cuh"
#include "summary.h"

using namespace b200;

namespace {

std::string g_create_error;

// ---- plug-in models (include/b200nuts_model.h): registry shared by all engines of the process
constexpr int MODEL_PLUGIN = B200NUTS_MODEL_PLUGIN_BASE;  // ModelData::kind of every plug-in model
std::mutex g_plugin_mutex;
std::vector<const b200nuts_model_vtable*> g_plugins;       // id = B200NUTS_MODEL_PLUGIN_BASE + index
const char* const kBuiltinNames[] = {"std_normal", "bernoulli", "eight_schools_nc", "logistic", "irt_2pl", "poisson_glm"};

const b200nuts_model_vtable* plugin_by_id(int id) {
  std::lock_guard<std::mutex> lk(g_plugin_mutex);
  const int i = id - B200NUTS_MODEL_PLUGIN_BASE;
  return (i >= 0 && i < (int)g_plugins.size()) ? g_plugins[i] : nullptr;
}

struct PluginVar {  // one data variable of a plug-in model, kept on the host until the model is created
  std::string name;
  int dtype;
  std::vector<int64_t> dims;
  std::vector<unsigned char> bytes;
};

struct ncclUniqueIdBlob {  // ncclUniqueId: 128 opaque bytes, passed by value
  char bytes[128];
};
struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, ncclUniqueIdBlob, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::string err;
};
NcclApi g_nccl;

#define CUDA_TRY(e, call)                                                                         \
  do {                                                                                            \
    cudaError_t err__ = (call);                                                                   \
    if (err__ != cudaSuccess) {                                                                   \
      (e)->err = std::string(#call) + ": " + cudaGetErrorString(err__);                           \
      return B200NUTS_SOFTWARE;                                                                   \
    }                                                                                             \
  } while (0)

// gradient stage of the lock-step engine for the small models (used for engine-equivalence tests)
__global__ void __launch_bounds__(128) grad_small_kernel(ModelData md, const double* Q, double* G, double* zV, const int* active,
                                                           int C, int ld) {
  const int c = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (c >= C) return;
  if (active && !active[c]) return;
  const int lane = threadIdx.x & 31;
  const double* q = Q + (size_t)c * ld;
  double* g = G + (size_t)c * ld;
  const double lp = lp_grad_small(md, q, g, lane);
  __syncwarp();
  for (int i = lane; i < md.D; i += 32) g[i] = -g[i];
  if (lane == 0) zV[c] = -lp;
}

// padded design matrix [1 | X | 0 ...]: the covariates were copied into columns 1 .. K (cudaMemcpy2D); column 0 and the padding here
__global__ void glm_pad_design_kernel(double* X, long long N, int K, int ldx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double* row = X + (size_t)i * ldx;
  row[0] = 1.0;
  for (int k = K + 1; k < ldx; k++) row[k] = 0.0;
}

__global__ void set_iter_end_kernel(ChainS* cs, int C, int add, int cap) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  int e = cs[c].iter_end + add;
  cs[c].iter_end = e > cap ? cap : e;
}

__global__ void init_metric_kernel(double* slab, size_t slab_stride, int ld, int D, int C, const double* user_im) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)C * D) return;
  const int c = (int)(idx / D), i = (int)(idx - (long long)c * D);
  slab[(size_t)c * slab_stride + (size_t)V_IM * ld + i] = user_im ? user_im[(size_t)c * D + i] : 1.0;
}

// rows of the draws buffer -> output rows with write_array applied: one thread per (chain, row)
__global__ void fetch_kernel(const double* draws, size_t draws_stride, int ncol, int kind, int D, int N, int K, int K_out,
                             int constrained, long long chain0, long long nchains, long long row0, long long nrows, double* out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nchains * nrows) return;
  const long long c = idx / nrows, r = idx - c * nrows;
  const double* src = draws + (size_t)(chain0 + c) * draws_stride + (size_t)(row0 + r) * ncol;
  const int kout = constrained ? K_out : D;
  double* dst = out + (size_t)idx * (7 + kout);
  for (int i = 0; i < 7; i++) dst[i] = src[i];
  if (constrained && kind < 0) return;  // plug-in model: its write_array is launched by the host
  if (constrained) model_write_array(kind, D, N, K, src + 7, dst + 7);
  else for (int i = 0; i < D; i++) dst[7 + i] = src[7 + i];
}

__global__ void gather_status_kernel(const ChainS* cs, int C, int* err, int* iter, long long* n_grad, double* eps, int* rows) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  err[c] = cs[c].err;
  iter[c] = cs[c].iter;
  n_grad[c] = cs[c].n_grad;
  eps[c] = cs[c].eps_nom;
  rows[c] = cs[c].draws_row;
}

// dense metric: mat[c][0] (column-major) <- identity or the user's [C][D][D] (row-major); and the way back
__global__ void init_dense_metric_kernel(double* mat, size_t mat_stride, int D, int C, const double* user_im) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long DD = (long long)D * D;
  if (idx >= (long long)C * DD) return;
  const int c = (int)(idx / DD);
  const int e = (int)(idx - (long long)c * DD), i = e / D, j = e - i * D;
  mat[(size_t)c * mat_stride + (size_t)j * D + i] = user_im ? user_im[idx] : (i == j ? 1.0 : 0.0);
}
__global__ void gather_dense_metric_kernel(const double* mat, size_t mat_stride, int D, int C, double* out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long DD = (long long)D * D;
  if (idx >= (long long)C * DD) return;
  const int c = (int)(idx / DD);
  const int e = (int)(idx - (long long)c * DD), i = e / D, j = e - i * D;
  out[idx] = mat[(size_t)c * mat_stride + (size_t)j * D + i];
}

__global__ void gather_metric_kernel(const double* slab, size_t slab_stride, int ld, int D, int C, double* out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)C * D) return;
  const int c = (int)(idx / D), i = (int)(idx - (long long)c * D);
  out[idx] = slab[(size_t)c * slab_stride + (size_t)V_IM * ld + i];
}

// log_prob with jacobian = 0: remove the log-Jacobian terms (and their derivatives) of the constraining transforms.
// in: zV = -lp, G = -grad (jacobian = 1).  One thread per point.
__global__ void remove_jacobian_kernel(int kind, int D, int N, int K, const double* Q, double* G, double* zV, int n, int ld) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const double* q = Q + (size_t)c * ld;
  double* g = G + (size_t)c * ld;
  if (kind == MODEL_BERNOULLI) {
    const double u = q[0], au = fabs(u);
    const double th = dev_inv_logit(u);
    zV[c] += -au - 2.0 * dev_log1p_exp(-au);
    g[0] += 1 - 2 * th;
  } else if (kind == MODEL_EIGHT_SCHOOLS) {
    zV[c] += q[1];
    g[1] += 1.0;
  } else if (kind == MODEL_IRT2PL) {
    const int I = N, J = K;
    double s = q[0] + q[1 + I] + q[3 + I + J];
    g[0] += 1.0;
    g[1 + I] += 1.0;
    g[3 + I + J] += 1.0;
    for (int j = 0; j < J; j++) {
      s += q[2 + I + j];
      g[2 + I + j] += 1.0;
    }
    zV[c] += s;
  }
}

}  // namespace

struct b200nuts_engine {
  b200nuts_config cfg;
  SamplerCfg scfg;
  ModelData md;
  std::string err;
  bool fused = false, data_ready = false, chains_ready = false;
  int C = 0, D = 0, ld = 0, K_out = 0;
  int sm_count = 148;
  // host staging of model data
  std::vector<double> h_yr, h_sigma;
  std::vector<int> h_yi;
  std::vector<int8_t> h_y8;
  int64_t x_rows = 0, x_cols = 0, y8_rows = 0, y8_cols = 0;
  int std_normal_D = 3;
  // plug-in model (cfg.model >= B200NUTS_MODEL_PLUGIN_BASE)
  const b200nuts_model_vtable* plug = nullptr;
  void* plug_model = nullptr;
  std::vector<PluginVar> plug_data;
  // device
  double *d_X = nullptr, *d_yr = nullptr, *d_sigma = nullptr;
  double* d_fetch = nullptr;   // staging buffer of b200nuts_fetch_draws
  size_t fetch_bytes = 0;
  int* d_yi = nullptr;
  signed char* d_y8 = nullptr;
  unsigned short* d_ypk = nullptr;
  DeviceChains dc{};
  double *d_user_q0 = nullptr, *d_user_im = nullptr;
  std::vector<double> h_user_q0, h_user_im;
  int *d_active = nullptr, *d_list = nullptr, *d_nactive = nullptr;  // d_nactive: ring of kCounterRing per-step counters
  int* h_nactive = nullptr;  // pinned
  void* d_glm_ws = nullptr;
  GlmWorkspace glm_ws;
  int glm_ws_N = -1, glm_ws_D = -1;  // data shape the workspace was sized for
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::vector<cudaEvent_t> grad_events;
  allreduce_fn ar = nullptr;
  void* ar_user = nullptr;
  void* nccl_comm = nullptr;  // ncclComm_t owned by the engine (b200nuts_nccl_init)
  int advance_cta = 1;        // long parameter vectors: one CTA per chain in the advance step (B200NUTS_ADVANCE_CTA=0: one warp per chain)
  int irt_solo = 48;          // IRT: at most this many requesting chains => the latency kernel finishes them (B200NUTS_IRT_SOLO; 0 = never)
  int irt_solo_cluster = 8;   // CTAs per chain of the latency kernel: 8 (cluster) or 1 (B200NUTS_IRT_SOLO_CLUSTER, or set after a failed cluster launch)
  double last_run_s = 0, last_grad_s = 0, last_solo_s = 0;
  int last_solo_chains = 0;
  long long last_launches = 0;
};

extern "C" {

void b200nuts_default_config(b200nuts_config* c) {
  std::memset(c, 0, sizeof *c);
  c->model = B200NUTS_STD_NORMAL;
  c->num_chains = 1;
  c->random_seed = 0;
  c->init_chain_id = 1;
  c->init_radius = 2.0;
  c->num_warmup = 1000;
  c->num_samples = 1000;
  c->num_thin = 1;
  c->save_warmup = 0;
  c->stepsize = 1.0;
  c->stepsize_jitter = 0.0;
  c->max_depth = 10;
  c->delta = 0.8;
  c->gamma = 0.05;
  c->kappa = 0.75;
  c->t0 = 10.0;
  c->init_buffer = 75;
  c->term_buffer = 50;
  c->window = 25;
  c->adapt_engaged = 1;
  c->engine = B200NUTS_ENGINE_AUTO;
  c->device = 0;
  c->store_draws = 1;
  c->row_shards = 1;
  c->row_shard_rank = 0;
  c->metric = 0;
  c->store_diagnostics = 0;
  c->static_hmc = 0;
  c->int_time = 6.28318530717958647692;
}

const char* b200nuts_last_error(const b200nuts_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int b200nuts_create(const b200nuts_config* cfg, b200nuts_engine** out) {
  *out = nullptr;
  // argument validation mirrors the reference's argument tree (src/cmdstan/arguments/arg_*.hpp is_valid)
  if (cfg->num_chains < 1) { g_create_error = "num_chains must be >= 1"; return B200NUTS_USAGE; }
  const b200nuts_model_vtable* plug = cfg->model >= B200NUTS_MODEL_PLUGIN_BASE ? plugin_by_id(cfg->model) : nullptr;
  if (cfg->model < 0 || (cfg->model > 5 && !plug)) { g_create_error = "unknown model"; return B200NUTS_USAGE; }
  if (plug && cfg->engine == B200NUTS_ENGINE_FUSED) { g_create_error = "plug-in models run in the lock-step engine"; return B200NUTS_CONFIG; }
  if (plug && cfg->row_shards > 1) { g_create_error = "row sharding is available for the built-in GLMs only"; return B200NUTS_CONFIG; }
  if (cfg->num_warmup < 0 || cfg->num_samples < 0 || cfg->num_thin < 1) { g_create_error = "num_warmup/num_samples must be >= 0 and thin >= 1"; return B200NUTS_USAGE; }
  if (cfg->max_depth < 1 || cfg->max_depth > kMaxDepthCap) { g_create_error = "max_depth must be in 1..16"; return B200NUTS_USAGE; }
  if (cfg->adapt_engaged && cfg->num_warmup == 0) {  // command.hpp:569-574
    g_create_error = "Model is adapting but num_warmup is 0: disable adaptation (adapt engaged=0) or set num_warmup > 0";
    return B200NUTS_USAGE;
  }
  if (!(cfg->delta > 0.1 - 1e-15 && cfg->delta < 1.0)) { g_create_error = "delta must be in [0.1, 1)"; return B200NUTS_USAGE; }
  if (!(cfg->stepsize > 0) || !(cfg->gamma > 0) || !(cfg->kappa > 0) || !(cfg->t0 > 0)) { g_create_error = "stepsize, gamma, kappa, t0 must be > 0"; return B200NUTS_USAGE; }
  if (cfg->stepsize_jitter < 0 || cfg->stepsize_jitter > 1) { g_create_error = "stepsize_jitter must be in [0, 1]"; return B200NUTS_USAGE; }
  if (cfg->init_radius < 0) { g_create_error = "init radius must be >= 0"; return B200NUTS_USAGE; }
  if (cfg->metric < 0 || cfg->metric > 2) { g_create_error = "metric must be 0 (diag_e), 1 (unit_e) or 2 (dense_e)"; return B200NUTS_USAGE; }
  if (cfg->static_hmc && !(cfg->int_time > 0)) { g_create_error = "int_time must be > 0"; return B200NUTS_USAGE; }
  if (cfg->static_hmc && cfg->store_diagnostics) { g_create_error = "store_diagnostics is available for the NUTS engine only"; return B200NUTS_CONFIG; }
  if (cfg->row_shards < 1 || cfg->row_shard_rank < 0 || cfg->row_shard_rank >= cfg->row_shards) { g_create_error = "bad row shard configuration"; return B200NUTS_USAGE; }
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0) {
    g_create_error = std::string("no CUDA device available (libb200nuts has no CPU path): ") + cudaGetErrorString(ce);
    return B200NUTS_SOFTWARE;
  }
  if (cfg->device < 0 || cfg->device >= ndev) { g_create_error = "device ordinal out of range"; return B200NUTS_USAGE; }
  auto* e = new b200nuts_engine();
  e->cfg = *cfg;
  e->C = cfg->num_chains;
  e->plug = plug;
  if (cudaSetDevice(cfg->device) != cudaSuccess || cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) {
    g_create_error = "cudaSetDevice / cudaStreamCreate failed";
    delete e;
    return B200NUTS_SOFTWARE;
  }
  cudaEventCreate(&e->ev0);
  cudaEventCreate(&e->ev1);
  cudaDeviceGetAttribute(&e->sm_count, cudaDevAttrMultiProcessorCount, cfg->device);
  std::memset(&e->md, 0, sizeof e->md);
  e->md.kind = plug ? MODEL_PLUGIN : cfg->model;
  *out = e;
  return B200NUTS_OK;
}

static void free_dev(void* p) { if (p) cudaFree(p); }

void b200nuts_destroy(b200nuts_engine* e) {
  if (!e) return;
  cudaSetDevice(e->cfg.device);
  if (e->plug && e->plug_model && e->plug->destroy) e->plug->destroy(e->plug_model);
  free_dev(e->d_X); free_dev(e->d_yr); free_dev(e->d_sigma); free_dev(e->d_yi); free_dev(e->d_y8); free_dev(e->d_ypk);
  free_dev(e->dc.cs); free_dev(e->dc.q); free_dev(e->dc.g); free_dev(e->dc.zV); free_dev(e->dc.slab); free_dev(e->dc.draws); free_dev(e->dc.mat);
  free_dev(e->d_user_q0); free_dev(e->d_user_im); free_dev(e->d_active); free_dev(e->d_list); free_dev(e->d_nactive); free_dev(e->d_glm_ws); free_dev(e->d_fetch);
  glm_i8_destroy(e->glm_ws.i8);
  if (e->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(e->nccl_comm);
  if (e->h_nactive) cudaFreeHost(e->h_nactive);
  for (auto ev : e->grad_events) cudaEventDestroy(ev);
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

static int64_t prod(const int64_t* dims, int nd) {
  int64_t p = 1;
  for (int i = 0; i < nd; i++) p *= dims[i];
  return p;
}

// A data variable was (re)set: the model dimensions may change, so everything sized from them is stale -- the chains (buffers,
// SamplerCfg D / ld) must be re-initialised (b200nuts_init_chains) and the GLM workspace is rebuilt for the new shapes.
static void data_changed(b200nuts_engine* e) {
  e->data_ready = false;
  e->chains_ready = false;
  e->glm_ws_N = e->glm_ws_D = -1;
}

// plug-in models: keep the variable as given; the model's create() validates and uploads it
static int plugin_set_data(b200nuts_engine* e, const char* name, int dtype, const void* host, const int64_t* dims, int nd) {
  static const size_t kSize[3] = {8, 4, 1};
  if (nd < 0 || nd > 8) { e->err = "set_data: 0 <= nd <= 8 required"; return B200NUTS_DATAERR; }
  PluginVar v;
  v.name = name;
  v.dtype = dtype;
  v.dims.assign(dims, dims + nd);
  const size_t bytes = (size_t)prod(dims, nd) * kSize[dtype];
  v.bytes.assign((const unsigned char*)host, (const unsigned char*)host + bytes);
  for (auto& o : e->plug_data)
    if (o.name == v.name) { o = std::move(v); data_changed(e); return B200NUTS_OK; }
  e->plug_data.push_back(std::move(v));
  data_changed(e);
  return B200NUTS_OK;
}

int b200nuts_set_data_f64(b200nuts_engine* e, const char* name, const double* host, const int64_t* dims, int nd) {
  if (e->plug) return plugin_set_data(e, name, B200NUTS_F64, host, dims, nd);
  const std::string n(name);
  const int64_t cnt = prod(dims, nd);
  if (n == "X" && nd == 2) {
    // straight to the device in the kernels' layout (row-major [N][ldx], column 0 = 1, zero padded): no host staging copy, no
    // padded host copy -- this upload is the host->device leg of the end-to-end time (80 MB at c3, 4 GB at c5).  The call returns
    // after the copy (synchronous), so the caller's buffer is consumed as the C ABI promises.
    if (dims[0] < 1 || dims[1] < 1 || dims[0] > 2147483647ll || ((dims[1] + 1 + 7) & ~7ll) > glm_max_dim()) {
      e->err = "glm: X must be [N >= 1][1 <= K], with K + 1 <= 176 in this build";
      return dims[1] < 1 || dims[0] < 1 ? B200NUTS_DATAERR : B200NUTS_CONFIG;
    }
    CUDA_TRY(e, cudaSetDevice(e->cfg.device));
    const int K = (int)dims[1], ldx = (K + 1 + 7) & ~7;
    if (!e->d_X || e->x_rows != dims[0] || e->x_cols != dims[1]) {  // a data set of the same shape is uploaded into the buffer in place
      free_dev(e->d_X);
      e->d_X = nullptr;
      CUDA_TRY(e, cudaMalloc(&e->d_X, sizeof(double) * (size_t)dims[0] * ldx));
    }
    CUDA_TRY(e, cudaMemcpy2DAsync(e->d_X + 1, sizeof(double) * ldx, host, sizeof(double) * K, sizeof(double) * K, (size_t)dims[0], cudaMemcpyHostToDevice, e->stream));
    glm_pad_design_kernel<<<(unsigned)((dims[0] + 255) / 256), 256, 0, e->stream>>>(e->d_X, dims[0], K, ldx);
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    e->x_rows = dims[0];
    e->x_cols = dims[1];
  } else if (n == "y" && e->cfg.model == B200NUTS_EIGHT_SCHOOLS_NC) {
    e->h_yr.assign(host, host + cnt);
  } else if (n == "sigma") {
    for (int64_t i = 0; i < cnt; i++)
      if (!(host[i] > 0)) { e->err = "sigma must be positive"; return B200NUTS_DATAERR; }
    e->h_sigma.assign(host, host + cnt);
  } else {
    e->err = "unknown f64 data variable '" + n + "' for this model";
    return B200NUTS_DATAERR;
  }
  data_changed(e);
  return B200NUTS_OK;
}

int b200nuts_set_data_i32(b200nuts_engine* e, const char* name, const int32_t* host, const int64_t* dims, int nd) {
  if (e->plug) return plugin_set_data(e, name, B200NUTS_I32, host, dims, nd);
  const std::string n(name);
  const int64_t cnt = prod(dims, nd);
  if (n == "y") {
    if (e->cfg.model == B200NUTS_BERNOULLI || e->cfg.model == B200NUTS_LOGISTIC_GLM) {
      for (int64_t i = 0; i < cnt; i++)
        if (host[i] < 0 || host[i] > 1) { e->err = "y must be in [0, 1]"; return B200NUTS_DATAERR; }
    } else if (e->cfg.model == B200NUTS_POISSON_GLM) {
      for (int64_t i = 0; i < cnt; i++)
        if (host[i] < 0) { e->err = "y must be nonnegative"; return B200NUTS_DATAERR; }
    }
    e->h_yi.assign(host, host + cnt);
  } else if (n == "D" && cnt == 1) {
    e->std_normal_D = host[0];
  } else {
    e->err = "unknown i32 data variable '" + n + "' for this model";
    return B200NUTS_DATAERR;
  }
  data_changed(e);
  return B200NUTS_OK;
}

int b200nuts_set_data_i8(b200nuts_engine* e, const char* name, const int8_t* host, const int64_t* dims, int nd) {
  if (e->plug) return plugin_set_data(e, name, B200NUTS_I8, host, dims, nd);
  const std::string n(name);
  if (n == "y" && nd == 2 && e->cfg.model == B200NUTS_IRT_2PL) {
    const int64_t cnt = prod(dims, nd);
    for (int64_t i = 0; i < cnt; i++)
      if (host[i] < 0 || host[i] > 1) { e->err = "y must be in [0, 1]"; return B200NUTS_DATAERR; }
    e->h_y8.assign(host, host + cnt);
    e->y8_rows = dims[0];
    e->y8_cols = dims[1];
    data_changed(e);
    return B200NUTS_OK;
  }
  e->err = "unknown i8 data variable '" + n + "' for this model";
  return B200NUTS_DATAERR;
}

// validates the data against the model, uploads it, fixes D / ld / engine kind
static int finalize_data(b200nuts_engine* e) {
  if (e->data_ready) return B200NUTS_OK;
  CUDA_TRY(e, cudaSetDevice(e->cfg.device));
  ModelData& m = e->md;
  if (e->plug) {  // new_model() of a plug-in: it validates + uploads its own data
    if (e->plug_model && e->plug->destroy) e->plug->destroy(e->plug_model);
    e->plug_model = nullptr;
    std::vector<b200nuts_data_view> views;
    for (auto& v : e->plug_data) views.push_back(b200nuts_data_view{v.name.c_str(), v.dtype, (int32_t)v.dims.size(), v.dims.data(), v.bytes.data()});
    int64_t D = 0, K_out = 0;
    char msg[512] = {0};
    const int rc = e->plug->create(views.data(), (int)views.size(), e->cfg.device, &e->plug_model, &D, &K_out, msg, (int)sizeof msg);
    if (rc != B200NUTS_OK) {
      e->err = std::string(e->plug->name) + ": " + (msg[0] ? msg : "model construction failed");
      e->plug_model = nullptr;
      return rc > 0 ? rc : B200NUTS_DATAERR;
    }
    if (D < 1 || K_out < 0 || (!e->plug->write_array && K_out != D)) { e->err = std::string(e->plug->name) + ": bad num_params_r / num_outputs"; return B200NUTS_CONFIG; }
    m.kind = MODEL_PLUGIN;
    m.D = (int)D;
    e->D = m.D;
    e->ld = (m.D + 3) & ~3;
    e->K_out = (int)K_out;
    e->fused = false;
    e->data_ready = true;
    return B200NUTS_OK;
  }
  const int kind = e->cfg.model;
  m.kind = kind;
  auto upload = [&](void** dptr, const void* src, size_t bytes) -> cudaError_t {
    if (*dptr) cudaFree(*dptr);
    *dptr = nullptr;
    cudaError_t r = cudaMalloc(dptr, bytes ? bytes : 8);
    if (r != cudaSuccess) return r;
    return cudaMemcpy(*dptr, src, bytes, cudaMemcpyHostToDevice);
  };
  switch (kind) {
    case B200NUTS_STD_NORMAL:
      if (e->std_normal_D < 1) { e->err = "D must be >= 1"; return B200NUTS_DATAERR; }
      m.D = e->std_normal_D;
      m.N = 0;
      break;
    case B200NUTS_BERNOULLI: {
      if (e->h_yi.empty()) { e->err = "bernoulli: data variable y missing"; return B200NUTS_NOINPUT; }
      m.D = 1;
      m.N = (int)e->h_yi.size();
      int s = 0;
      for (int v : e->h_yi) s += v;
      m.sum_y = s;
      break;
    }
    case B200NUTS_EIGHT_SCHOOLS_NC: {
      if (e->h_yr.empty() || e->h_yr.size() != e->h_sigma.size()) { e->err = "eight_schools: y and sigma must be given with equal sizes"; return B200NUTS_NOINPUT; }
      m.N = (int)e->h_yr.size();
      m.D = m.N + 2;
      CUDA_TRY(e, upload((void**)&e->d_yr, e->h_yr.data(), e->h_yr.size() * 8));
      CUDA_TRY(e, upload((void**)&e->d_sigma, e->h_sigma.data(), e->h_sigma.size() * 8));
      m.yr = e->d_yr;
      m.sigma = e->d_sigma;
      break;
    }
    case B200NUTS_LOGISTIC_GLM:
    case B200NUTS_POISSON_GLM: {
      if (!e->d_X || (int64_t)e->h_yi.size() != e->x_rows) { e->err = "glm: X[N][K] and y[N] must be given"; return B200NUTS_NOINPUT; }
      m.N = (int)e->x_rows;
      m.K = (int)e->x_cols;
      m.D = m.K + 1;
      if (((m.D + 7) & ~7) > glm_max_dim()) { e->err = "glm: too many covariates for this build (D <= 176)"; return B200NUTS_CONFIG; }
      m.ldx = (m.D + 7) & ~7;  // the padded design matrix [1 | X | 0] is already on the device (b200nuts_set_data_f64)
      CUDA_TRY(e, upload((void**)&e->d_yi, e->h_yi.data(), e->h_yi.size() * 4));
      m.X = e->d_X;
      m.yi = e->d_yi;
      break;
    }
    case B200NUTS_IRT_2PL: {
      if (e->h_y8.empty()) { e->err = "irt_2pl: y[I][J] missing"; return B200NUTS_NOINPUT; }
      m.N = (int)e->y8_rows;
      m.K = (int)e->y8_cols;
      if (m.K > irt_max_items()) { e->err = "irt_2pl: too many items for this build (J <= 128)"; return B200NUTS_CONFIG; }
      m.D = 4 + m.N + 2 * m.K;
      CUDA_TRY(e, upload((void**)&e->d_y8, e->h_y8.data(), e->h_y8.size()));
      m.y8 = e->d_y8;
      {
        std::vector<unsigned short> pk((size_t)m.N * 8);
        irt_pack_responses(e->h_y8.data(), m.N, m.K, pk.data());
        CUDA_TRY(e, upload((void**)&e->d_ypk, pk.data(), pk.size() * sizeof(unsigned short)));
        m.ypk = e->d_ypk;
      }
      break;
    }
  }
  e->D = m.D;
  e->ld = (m.D + 3) & ~3;
  e->K_out = model_num_outputs(kind, m.D, m.N);
  const bool small = kind <= B200NUTS_EIGHT_SCHOOLS_NC;
  if (e->cfg.engine == B200NUTS_ENGINE_FUSED && !small) { e->err = "the fused engine supports models 0-2 only"; return B200NUTS_CONFIG; }
  e->fused = e->cfg.engine == B200NUTS_ENGINE_FUSED || (e->cfg.engine == B200NUTS_ENGINE_AUTO && small);
  if (e->cfg.row_shards > 1 && !(kind == B200NUTS_LOGISTIC_GLM || kind == B200NUTS_POISSON_GLM)) { e->err = "row sharding is available for the GLMs only"; return B200NUTS_CONFIG; }
  e->data_ready = true;
  return B200NUTS_OK;
}

int b200nuts_num_params(const b200nuts_engine* ce, int64_t* nu, int64_t* no) {
  auto* e = const_cast<b200nuts_engine*>(ce);
  int rc = finalize_data(e);
  if (rc) return rc;
  if (nu) *nu = e->D;
  if (no) *no = e->K_out;
  return B200NUTS_OK;
}

int b200nuts_set_inits(b200nuts_engine* e, const double* q0) {
  int rc = finalize_data(e);
  if (rc) return rc;
  if (q0) e->h_user_q0.assign(q0, q0 + (size_t)e->C * e->D);
  else e->h_user_q0.clear();
  return B200NUTS_OK;
}

int b200nuts_set_inv_metric(b200nuts_engine* e, const double* im) {
  int rc = finalize_data(e);
  if (rc) return rc;
  if (im && e->cfg.metric == 1) { e->err = "unit_e has no free metric parameters"; return B200NUTS_USAGE; }
  if (im && e->cfg.metric == 2) {
    // validate_dense_inv_metric (S/services/util/validate_dense_inv_metric.hpp): check_pos_definite = symmetric + LDLT positive
    const size_t D = (size_t)e->D;
    std::vector<double> L(D * D);
    for (int c = 0; c < e->C; c++) {
      const double* A = im + (size_t)c * D * D;
      bool ok = true;
      for (size_t i = 0; i < D && ok; i++)
        for (size_t j = 0; j < i; j++)
          if (!(std::fabs(A[i * D + j] - A[j * D + i]) <= 1e-8)) { ok = false; break; }  // check_symmetric: CONSTRAINT_TOLERANCE 1e-8
      for (size_t j = 0; j < D && ok; j++) {  // Cholesky of the lower triangle
        double d = A[j * D + j];
        for (size_t k = 0; k < j; k++) d -= L[j * D + k] * L[j * D + k];
        if (!(d > 0) || !std::isfinite(d)) { ok = false; break; }
        L[j * D + j] = std::sqrt(d);
        for (size_t i = j + 1; i < D; i++) {
          double v = A[i * D + j];
          for (size_t k = 0; k < j; k++) v -= L[i * D + k] * L[j * D + k];
          L[i * D + j] = v / L[j * D + j];
        }
      }
      if (!ok) { e->err = "Inverse Euclidean metric not positive definite."; return B200NUTS_CONFIG; }
    }
    e->h_user_im.assign(im, im + (size_t)e->C * D * D);
  } else if (im) {
    for (size_t i = 0; i < (size_t)e->C * e->D; i++)
      if (!(im[i] > 0) || !std::isfinite(im[i])) { e->err = "Inverse Euclidean metric not positive definite."; return B200NUTS_CONFIG; }  // validate_diag_inv_metric
    e->h_user_im.assign(im, im + (size_t)e->C * e->D);
  } else {
    e->h_user_im.clear();
  }
  return B200NUTS_OK;
}

// ---- NCCL, loaded at run time (the library has no link-time dependency on NCCL; single-GPU users never touch it) --------
namespace {

bool load_nccl(std::string* err) {
  if (g_nccl.handle) return true;
  std::vector<std::string> names;
  if (const char* p = getenv("B200NUTS_NCCL_LIB")) names.push_back(p);
  names.push_back("libnccl.so.2");
  names.push_back("libnccl.so");
  for (const std::string& n : names) {
    g_nccl.handle = dlopen(n.c_str(), RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.handle) break;
  }
  if (!g_nccl.handle) {
    *err = "cannot load NCCL (set B200NUTS_NCCL_LIB to the path of libnccl.so.2; the torch-bundled one lives in site-packages/nvidia/nccl/lib)";
    return false;
  }
  g_nccl.GetUniqueId = (int (*)(void*))dlsym(g_nccl.handle, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(void**, int, ncclUniqueIdBlob, int))dlsym(g_nccl.handle, "ncclCommInitRank");
  g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(g_nccl.handle, "ncclAllReduce");
  g_nccl.CommDestroy = (int (*)(void*))dlsym(g_nccl.handle, "ncclCommDestroy");
  g_nccl.GetErrorString = (const char* (*)(int))dlsym(g_nccl.handle, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) {
    *err = "the NCCL library lacks ncclGetUniqueId / ncclCommInitRank / ncclAllReduce / ncclCommDestroy";
    return false;
  }
  return true;
}

// the per-leapfrog exchange step of row sharding: sum the (lp, gradient) partials of all row shards, in place
int nccl_allreduce_hook(void* device_buf, int64_t count, void* cuda_stream, void* user) {
  const int kDouble = 8, kSum = 0;  // ncclFloat64, ncclSum
  return g_nccl.AllReduce(device_buf, device_buf, (size_t)count, kDouble, kSum, user, (cudaStream_t)cuda_stream);
}
}  // namespace

int b200nuts_nccl_get_unique_id(void* id128) {
  std::string err;
  if (!load_nccl(&err)) {
    g_create_error = err;
    return B200NUTS_SOFTWARE;
  }
  int rc = g_nccl.GetUniqueId(id128);
  if (rc) {
    g_create_error = std::string("ncclGetUniqueId: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error");
    return B200NUTS_SOFTWARE;
  }
  return B200NUTS_OK;
}

int b200nuts_nccl_init(b200nuts_engine* e, const void* id128, int nranks, int rank) {
  if (!load_nccl(&e->err)) return B200NUTS_SOFTWARE;
  if (nranks != e->cfg.row_shards || rank != e->cfg.row_shard_rank) {
    e->err = "b200nuts_nccl_init: nranks / rank must equal the engine's row_shards / row_shard_rank";
    return B200NUTS_USAGE;
  }
  CUDA_TRY(e, cudaSetDevice(e->cfg.device));
  ncclUniqueIdBlob id;
  std::memcpy(id.bytes, id128, sizeof id.bytes);
  void* comm = nullptr;
  int rc = g_nccl.CommInitRank(&comm, nranks, id, rank);
  if (rc) {
    e->err = std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error");
    return B200NUTS_SOFTWARE;
  }
  e->nccl_comm = comm;
  e->ar = nccl_allreduce_hook;
  e->ar_user = comm;
  return B200NUTS_OK;
}

int b200nuts_measure_allreduce(b200nuts_engine* e, int iters, double* usec, int64_t* payload_bytes) {
  if (!e->ar || e->cfg.row_shards <= 1) { e->err = "measure_allreduce: the engine is not row-sharded (no all-reduce hook installed)"; return B200NUTS_USAGE; }
  if (!e->d_glm_ws) { e->err = "measure_allreduce: b200nuts_init_chains must be called first"; return B200NUTS_USAGE; }
  if (iters < 1) iters = 1;
  CUDA_TRY(e, cudaSetDevice(e->cfg.device));
  const long long count = (long long)e->glm_ws.C * (1 + e->glm_ws.Dp);
  CUDA_TRY(e, cudaMemsetAsync(e->glm_ws.R, 0, sizeof(double) * (size_t)count, e->stream));
  for (int w = 0; w < 3; w++)
    if (e->ar(e->glm_ws.R, count, (void*)e->stream, e->ar_user)) { e->err = "measure_allreduce: the all-reduce hook failed"; return B200NUTS_SOFTWARE; }
  CUDA_TRY(e, cudaEventRecord(e->ev0, e->stream));
  for (int i = 0; i < iters; i++)
    if (e->ar(e->glm_ws.R, count, (void*)e->stream, e->ar_user)) { e->err = "measure_allreduce: the all-reduce hook failed"; return B200NUTS_SOFTWARE; }
  CUDA_TRY(e, cudaEventRecord(e->ev1, e->stream));
  CUDA_TRY(e, cudaStreamSynchronize(e->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, e->ev0, e->ev1);
  if (usec) *usec = 1e3 * ms / iters;
  if (payload_bytes) *payload_bytes = (int64_t)count * 8;
  return B200NUTS_OK;
}

int b200nuts_set_allreduce(b200nuts_engine* e, b200nuts_allreduce_fn fn, void* user) {
  e->ar = fn;
  e->ar_user = user;
  return B200NUTS_OK;
}

static int ensure_glm_ws(b200nuts_engine* e, int C) {
  if (!(e->md.kind == MODEL_LOGISTIC || e->md.kind == MODEL_POISSON)) return B200NUTS_OK;
  if (e->d_glm_ws && e->glm_ws.C >= C && e->glm_ws_N == e->md.N && e->glm_ws_D == e->md.D) return B200NUTS_OK;
  free_dev(e->d_glm_ws);
  e->d_glm_ws = nullptr;
  glm_i8_destroy(e->glm_ws.i8);
  e->glm_ws.i8 = nullptr;
  size_t bytes = glm_workspace_bytes(e->md.N, e->md.D, C, e->sm_count, &e->glm_ws);
  CUDA_TRY(e, cudaMalloc(&e->d_glm_ws, bytes));
  glm_workspace_bind(&e->glm_ws, e->d_glm_ws);
  e->glm_ws_N = e->md.N;  // P is a function of N, Dp of D: a workspace is only reusable for the shapes it was sized for
  e->glm_ws_D = e->md.D;
  // the all-reduce payload of a row-sharded run covers the rows of every chain, requesting or not: never uninitialised memory
  CUDA_TRY(e, cudaMemsetAsync(e->glm_ws.R, 0, sizeof(double) * (size_t)e->glm_ws.C * (1 + e->glm_ws.Dp), e->stream));
  const char* v = getenv("B200NUTS_GLM_VARIANT");
  // default: the INT8 digit-plane kernel on tcgen05 for the logistic link (r02: 2x the FP64 tensor kernel), the FP64 tensor kernel v2 else
  e->glm_ws.variant = v ? atoi(v) : (e->md.kind == MODEL_LOGISTIC ? 7 : 1);
  if (e->glm_ws.variant == 7) {  // INT8 digit planes on tcgen05; where it does not apply (nullptr) the FP64 tensor kernel v2 serves
    e->glm_ws.i8 = glm_i8_create(e->md, e->glm_ws.Dp, e->glm_ws.C, e->sm_count, e->glm_ws.P, e->stream);
    if (!e->glm_ws.i8) e->glm_ws.variant = 1;
  }
  return B200NUTS_OK;
}

int b200nuts_init_chains(b200nuts_engine* e) {
  int rc = finalize_data(e);
  if (rc) return rc;
  CUDA_TRY(e, cudaSetDevice(e->cfg.device));
  const b200nuts_config& c = e->cfg;
  SamplerCfg& s = e->scfg;
  std::memset(&s, 0, sizeof s);
  s.D = e->D;
  s.ld = e->ld;
  s.max_depth = c.max_depth;
  s.num_warmup = c.num_warmup;
  s.num_samples = c.num_samples;
  s.num_thin = c.num_thin;
  s.save_warmup = c.save_warmup;
  s.adapt_engaged = c.adapt_engaged;
  s.user_inits = !e->h_user_q0.empty();
  s.stepsize = c.stepsize;
  s.stepsize_jitter = (c.stepsize_jitter > 0 && c.stepsize_jitter < 1) ? c.stepsize_jitter : 0.0;  // base_hmc::set_stepsize_jitter
  s.delta = c.delta;
  s.gamma = c.gamma;
  s.kappa = c.kappa;
  s.t0 = c.t0;
  s.init_radius = c.init_radius;
  // unit_e: no variance windows at all (adapt_unit_e_nuts::transition only learns the step size)
  s.dense = c.metric == 2;
  resolve_window(&s, (c.adapt_engaged && c.metric != 1) ? (unsigned)c.num_warmup : 0u, c.init_buffer, c.term_buffer, c.window);
  s.diag = c.store_diagnostics && c.store_draws ? 1 : 0;
  s.static_hmc = c.static_hmc ? 1 : 0;
  s.int_time = c.int_time;
  s.ncol = 7 + e->D * (s.diag ? 3 : 1);
  s.rows_per_chain = c.store_draws ? (c.save_warmup ? (c.num_warmup + c.num_thin - 1) / c.num_thin : 0) + (c.num_samples + c.num_thin - 1) / c.num_thin : 0;

  const int C = e->C;
  DeviceChains& d = e->dc;
  free_dev(d.cs); free_dev(d.q); free_dev(d.g); free_dev(d.zV); free_dev(d.slab); free_dev(d.draws); free_dev(d.mat);
  free_dev(e->d_user_q0); free_dev(e->d_user_im); free_dev(e->d_active); free_dev(e->d_list); free_dev(e->d_nactive);
  d = DeviceChains{};
  e->d_user_q0 = e->d_user_im = nullptr;
  d.slab_stride = (size_t)num_vec_slots(c.max_depth, s.dense, s.diag) * e->ld;
  d.mat_stride = s.dense ? 3 * (size_t)e->D * e->D : 0;
  d.draws_stride = (size_t)s.rows_per_chain * s.ncol;
  CUDA_TRY(e, cudaMalloc(&d.cs, sizeof(ChainS) * C));
  CUDA_TRY(e, cudaMalloc(&d.q, sizeof(double) * C * e->ld));
  CUDA_TRY(e, cudaMalloc(&d.g, sizeof(double) * C * e->ld));
  CUDA_TRY(e, cudaMalloc(&d.zV, sizeof(double) * C));
  CUDA_TRY(e, cudaMalloc(&d.slab, sizeof(double) * C * d.slab_stride));
  CUDA_TRY(e, cudaMemset(d.q, 0, sizeof(double) * C * e->ld));
  CUDA_TRY(e, cudaMemset(d.g, 0, sizeof(double) * C * e->ld));
  CUDA_TRY(e, cudaMemset(d.zV, 0, sizeof(double) * C));
  CUDA_TRY(e, cudaMemset(d.slab, 0, sizeof(double) * C * d.slab_stride));
  if (s.dense) {
    cudaError_t r = cudaMalloc(&d.mat, sizeof(double) * C * d.mat_stride);
    if (r != cudaSuccess) {
      e->err = "cannot allocate the dense metric (" + std::to_string(sizeof(double) * C * d.mat_stride >> 20) + " MiB): " + cudaGetErrorString(r);
      d.mat = nullptr;
      return B200NUTS_SOFTWARE;
    }
    CUDA_TRY(e, cudaMemset(d.mat, 0, sizeof(double) * C * d.mat_stride));
  }
  if (d.draws_stride) {
    cudaError_t r = cudaMalloc(&d.draws, sizeof(double) * C * d.draws_stride);
    if (r != cudaSuccess) {
      e->err = "cannot allocate the draws buffer (" + std::to_string(sizeof(double) * C * d.draws_stride >> 20) + " MiB): " + cudaGetErrorString(r) + "; use store_draws=0, thin, or fewer chains per GPU";
      d.draws = nullptr;
      return B200NUTS_SOFTWARE;
    }
  }
  CUDA_TRY(e, cudaMalloc(&e->d_active, sizeof(int) * C));
  CUDA_TRY(e, cudaMalloc(&e->d_list, sizeof(int) * C));
  CUDA_TRY(e, cudaMalloc(&e->d_nactive, sizeof(int) * kCounterRing));
  if (!e->h_nactive) CUDA_TRY(e, cudaMallocHost(&e->h_nactive, sizeof(int)));
  // per-chain scalars + seeded streams (host), then upload
  std::vector<ChainS> hs(C);
  for (int i = 0; i < C; i++) {
    std::memset(&hs[i], 0, sizeof(ChainS));
    chain_init_scalars(&hs[i], s);
    seed_stream(hs[i].rng, c.random_seed, c.init_chain_id + (uint32_t)i);
  }
  CUDA_TRY(e, cudaMemcpy(d.cs, hs.data(), sizeof(ChainS) * C, cudaMemcpyHostToDevice));
  if (!e->h_user_q0.empty()) {
    std::vector<double> padded((size_t)C * e->ld, 0.0);
    for (int i = 0; i < C; i++) std::memcpy(padded.data() + (size_t)i * e->ld, e->h_user_q0.data() + (size_t)i * e->D, sizeof(double) * e->D);
    CUDA_TRY(e, cudaMalloc(&e->d_user_q0, sizeof(double) * padded.size()));
    CUDA_TRY(e, cudaMemcpy(e->d_user_q0, padded.data(), sizeof(double) * padded.size(), cudaMemcpyHostToDevice));
    d.user_q0 = e->d_user_q0;
  }
  if (!e->h_user_im.empty()) {
    CUDA_TRY(e, cudaMalloc(&e->d_user_im, sizeof(double) * e->h_user_im.size()));
    CUDA_TRY(e, cudaMemcpy(e->d_user_im, e->h_user_im.data(), sizeof(double) * e->h_user_im.size(), cudaMemcpyHostToDevice));
  }
  if (s.dense) {
    const long long tot = (long long)C * e->D * e->D;
    if (!e->h_user_im.empty() && (long long)e->h_user_im.size() != tot) { e->err = "dense_e: the inverse metric must be [num_chains][D][D]"; return B200NUTS_USAGE; }
    init_dense_metric_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, e->stream>>>(d.mat, d.mat_stride, e->D, C, e->d_user_im);
  } else {
    const long long tot = (long long)C * e->D;
    if (!e->h_user_im.empty() && (long long)e->h_user_im.size() != tot) { e->err = "diag_e: the inverse metric must be [num_chains][D]"; return B200NUTS_USAGE; }
    init_metric_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, e->stream>>>(d.slab, d.slab_stride, e->ld, e->D, C, e->d_user_im);
  }
  rc = ensure_glm_ws(e, C);
  if (rc) return rc;
  if (const char* t = getenv("B200NUTS_ADVANCE_CTA")) e->advance_cta = atoi(t) != 0;
  if (const char* t = getenv("B200NUTS_IRT_SOLO")) e->irt_solo = atoi(t);
  if (const char* t = getenv("B200NUTS_IRT_SOLO_CLUSTER")) e->irt_solo_cluster = atoi(t) == 1 ? 1 : 8;
  CUDA_TRY(e, cudaStreamSynchronize(e->stream));
  CUDA_TRY(e, cudaGetLastError());
  e->chains_ready = true;
  return B200NUTS_OK;
}

// gradient stage of the lock-step engine; returns the number of kernels launched
static int launch_gradient(b200nuts_engine* e, const GradBatch& b, int jacobian = 1) {
  if (e->plug) {
    const int n = e->plug->log_prob_grad(e->plug_model, b.Q, b.G, b.zV, b.list, b.n_list, b.C, b.ld, jacobian, (void*)e->stream);
    return n > 0 ? n : 0;
  }
  switch (e->md.kind) {
    case MODEL_LOGISTIC:
    case MODEL_POISSON:
      return launch_glm_grad(e->md, b, e->glm_ws, e->cfg.row_shards > 1 ? e->ar : nullptr, e->ar_user, true, e->stream);
    case MODEL_IRT2PL:
      return launch_irt_grad(e->md, b, e->stream);
    default:
      grad_small_kernel<<<(b.C + 3) / 4, 128, 0, e->stream>>>(e->md, b.Q, b.G, b.zV, b.active, b.C, b.ld);
      return 1;
  }
}

int b200nuts_run(b200nuts_engine* e, int n_iter) {
  if (!e->chains_ready) { e->err = "b200nuts_init_chains must be called first"; return B200NUTS_USAGE; }
  if (e->cfg.row_shards > 1 && !e->ar) { e->err = "row_shards > 1 needs b200nuts_set_allreduce"; return B200NUTS_USAGE; }
  CUDA_TRY(e, cudaSetDevice(e->cfg.device));
  const int C = e->C;
  const int total = e->cfg.num_warmup + e->cfg.num_samples;
  long long launches = 0;
  double grad_ms = 0;
  e->last_solo_s = 0;
  e->last_solo_chains = 0;
  CUDA_TRY(e, cudaEventRecord(e->ev0, e->stream));
  set_iter_end_kernel<<<(C + 255) / 256, 256, 0, e->stream>>>(e->dc.cs, C, n_iter, total);
  launches++;
  const int blocks = (C + kChainsPerBlock - 1) / kChainsPerBlock;
  if (e->fused) {
    // state in shared memory when a wave of ceil(C / #SM) chains per SM fits (<= 16 KB per chain); else global memory
    const int chain_words = kChainWords + 2 * e->ld + 2 + (int)e->dc.slab_stride;
    const size_t chain_bytes = (size_t)chain_words * 8;
    const int per_sm = (C + e->sm_count - 1) / e->sm_count;
    if (chain_bytes <= 16 * 1024) {
      // warps per block: spread the per-SM wave over blocks of <= 8 warps so that several blocks share an SM
      int wpb = per_sm < kFusedMaxWarps ? (per_sm < 1 ? 1 : per_sm) : kFusedMaxWarps;
      while ((size_t)wpb * chain_bytes > 56 * 1024 && wpb > 1) wpb--;
      const size_t smem = (size_t)wpb * chain_bytes;
      if (e->scfg.dense || e->scfg.diag || e->scfg.static_hmc) {
        CUDA_TRY(e, launch_fused_dense(e->scfg, e->md, e->dc, C, true, wpb, chain_words, smem, e->stream));
      } else {
        auto k = fused_chain_kernel<true, 0>;
        CUDA_TRY(e, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<(C + wpb - 1) / wpb, 32 * wpb, smem, e->stream>>>(e->scfg, e->md, e->dc, C, wpb, chain_words);
      }
    } else if (e->scfg.dense || e->scfg.diag || e->scfg.static_hmc) {
      CUDA_TRY(e, launch_fused_dense(e->scfg, e->md, e->dc, C, false, kChainsPerBlock, 0, 0, e->stream));
    } else {
      fused_chain_kernel<false, 0><<<blocks, 32 * kChainsPerBlock, 0, e->stream>>>(e->scfg, e->md, e->dc, C, kChainsPerBlock, 0);
    }
    launches++;
  } else {
    GradBatch b{e->dc.q, e->dc.g, e->dc.zV, e->d_active, C, e->ld};
    b.list = e->d_list;
    const int kCheck = kCounterRing;
    // two CUDA events bracket every gradient stage so that the dominant kernel family is timed live (bench.py roofline)
    size_t ev_used = 0;
    auto next_event = [&]() -> cudaEvent_t {
      if (ev_used == e->grad_events.size()) {
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        e->grad_events.push_back(ev);
      }
      return e->grad_events[ev_used++];
    };
    for (long long step = 0;; step++) {
      const bool check = (step % kCheck) == 0;
      int* counter = e->d_nactive + (step % kCounterRing);  // one counter per step of the ring, zeroed together
      if (check) CUDA_TRY(e, cudaMemsetAsync(e->d_nactive, 0, sizeof(int) * kCounterRing, e->stream));
      if (e->scfg.dense || e->scfg.diag || e->scfg.static_hmc) launch_advance_dense(e->scfg, e->dc, C, e->d_active, e->d_list, counter, e->stream);
      else if (e->advance_cta && e->D >= kAdvanceCtaD && e->ld <= kAdvanceCtaMaxLd) {  // two kernels: leaf path per CTA, the rest per warp
        CUDA_TRY(e, launch_advance_cta(e->scfg, e->dc, C, e->d_active, e->d_list, counter, e->stream));
        launches++;
      }
      else if (e->D >= kAdvanceBigD) launch_advance_big(e->scfg, e->dc, C, e->d_active, e->d_list, counter, e->stream);
      else advance_kernel<0><<<blocks, 32 * kChainsPerBlock, 0, e->stream>>>(e->scfg, e->dc, C, e->d_active, e->d_list, counter);
      launches++;
      b.n_list = counter;
      if (check) {
        CUDA_TRY(e, cudaMemcpyAsync(e->h_nactive, counter, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
        CUDA_TRY(e, cudaStreamSynchronize(e->stream));
        // harvest finished event pairs
        for (size_t i = 0; i + 1 < ev_used; i += 2) {
          float ms = 0;
          cudaEventElapsedTime(&ms, e->grad_events[i], e->grad_events[i + 1]);
          grad_ms += ms;
        }
        ev_used = 0;
        if (*e->h_nactive == 0) break;
        b.hint_active = *e->h_nactive;
        if (e->irt_solo > 0 && e->md.kind == MODEL_IRT2PL && !e->plug && !e->scfg.dense && !e->scfg.diag && !e->scfg.static_hmc &&
            *e->h_nactive <= e->irt_solo && irt_solo_supported(e->md)) {
          // only stragglers are left: each gets a cluster of eight SMs that runs it to the end of this call on the device (irt.cu)
          cudaEvent_t a = next_event(), bb = next_event();
          cudaEventRecord(a, e->stream);
          cudaError_t le = launch_irt_solo(e->scfg, e->dc, e->md, e->d_list, counter, *e->h_nactive, e->irt_solo_cluster, e->stream);
          if (le != cudaSuccess && e->irt_solo_cluster != 1) {  // no cluster launch on this device / configuration: one CTA per chain
            cudaGetLastError();
            e->irt_solo_cluster = 1;
            le = launch_irt_solo(e->scfg, e->dc, e->md, e->d_list, counter, *e->h_nactive, 1, e->stream);
          }
          if (le != cudaSuccess) { e->err = std::string("launch of the IRT latency kernel: ") + cudaGetErrorString(le); return B200NUTS_SOFTWARE; }
          cudaEventRecord(bb, e->stream);
          launches++;
          CUDA_TRY(e, cudaStreamSynchronize(e->stream));
          float solo_ms = 0;
          cudaEventElapsedTime(&solo_ms, a, bb);
          grad_ms += solo_ms;  // gradient + the stragglers' own advance work
          e->last_solo_s = solo_ms * 1e-3;
          e->last_solo_chains = *e->h_nactive;
          if (getenv("B200NUTS_VERBOSE"))
            std::fprintf(stderr, "[b200nuts] latency kernel: %d chains handed over after %lld lock-steps, %.3f s (cluster of %d)\n", *e->h_nactive, step,
                         e->last_solo_s, e->irt_solo_cluster);
          break;
        }
      }
      cudaEvent_t a = next_event(), bb = next_event();
      cudaEventRecord(a, e->stream);
      const int nl = launch_gradient(e, b);
      cudaEventRecord(bb, e->stream);
      if (nl < 0) {  // the exchange step of a row-sharded run failed: stop here, the chains' state is no longer meaningful
        cudaStreamSynchronize(e->stream);
        e->chains_ready = false;
        e->err = "all-reduce of the (log density, gradient) partials failed with code " + std::to_string(-nl) + "; the run was stopped";
        return B200NUTS_SOFTWARE;
      }
      launches += nl;
    }
  }
  CUDA_TRY(e, cudaEventRecord(e->ev1, e->stream));
  CUDA_TRY(e, cudaStreamSynchronize(e->stream));
  CUDA_TRY(e, cudaGetLastError());
  float ms = 0;
  cudaEventElapsedTime(&ms, e->ev0, e->ev1);
  e->last_run_s = ms * 1e-3;
  e->last_grad_s = grad_ms * 1e-3;
  e->last_launches = launches;
  return B200NUTS_OK;
}

int b200nuts_get_timing(b200nuts_engine* e, double* run_s, int64_t* launches, double* grad_s) {
  if (run_s) *run_s = e->last_run_s;
  if (launches) *launches = e->last_launches;
  if (grad_s) *grad_s = e->last_grad_s;
  return B200NUTS_OK;
}

int b200nuts_log_prob_grad(b200nuts_engine* e, const double* q, int64_t n, int jacobian, double* lp, double* grad) {
  int rc = finalize_data(e);
  if (rc) return rc;
  if (e->cfg.row_shards > 1 && !e->ar) { e->err = "row_shards > 1 needs b200nuts_set_allreduce"; return B200NUTS_USAGE; }
  CUDA_TRY(e, cudaSetDevice(e->cfg.device));
  const int D = e->D, ld = e->ld;
  // evaluate in batches so that temporary memory stays bounded
  const int64_t batch = std::min<int64_t>(n, 4096);
  double *dQ = nullptr, *dG = nullptr, *dV = nullptr;
  CUDA_TRY(e, cudaMalloc(&dQ, sizeof(double) * batch * ld));
  CUDA_TRY(e, cudaMalloc(&dG, sizeof(double) * batch * ld));
  CUDA_TRY(e, cudaMalloc(&dV, sizeof(double) * batch));
  rc = ensure_glm_ws(e, (int)std::max<int64_t>(batch, e->C));
  if (rc) { cudaFree(dQ); cudaFree(dG); cudaFree(dV); return rc; }
  std::vector<double> hq((size_t)batch * ld, 0.0), hg((size_t)batch * ld), hv(batch);
  for (int64_t off = 0; off < n; off += batch) {
    const int m = (int)std::min<int64_t>(batch, n - off);
    for (int i = 0; i < m; i++) std::memcpy(hq.data() + (size_t)i * ld, q + (size_t)(off + i) * D, sizeof(double) * D);
    cudaMemcpyAsync(dQ, hq.data(), sizeof(double) * m * ld, cudaMemcpyHostToDevice, e->stream);
    GradBatch b{dQ, dG, dV, nullptr, m, ld};
    if (launch_gradient(e, b, jacobian) < 0) {
      cudaStreamSynchronize(e->stream);
      cudaFree(dQ); cudaFree(dG); cudaFree(dV);
      e->err = "log_prob_grad: the all-reduce of the row-shard partials failed";
      return B200NUTS_SOFTWARE;
    }
    if (!jacobian && !e->plug) remove_jacobian_kernel<<<(m + 127) / 128, 128, 0, e->stream>>>(e->md.kind, D, e->md.N, e->md.K, dQ, dG, dV, m, ld);
    cudaMemcpyAsync(hg.data(), dG, sizeof(double) * m * ld, cudaMemcpyDeviceToHost, e->stream);
    cudaMemcpyAsync(hv.data(), dV, sizeof(double) * m, cudaMemcpyDeviceToHost, e->stream);
    cudaError_t r = cudaStreamSynchronize(e->stream);
    if (r == cudaSuccess) r = cudaGetLastError();
    if (r != cudaSuccess) {
      e->err = std::string("log_prob_grad: ") + cudaGetErrorString(r);
      cudaFree(dQ); cudaFree(dG); cudaFree(dV);
      return B200NUTS_SOFTWARE;
    }
    for (int i = 0; i < m; i++) {
      lp[off + i] = -hv[i];
      for (int k = 0; k < D; k++) grad[(size_t)(off + i) * D + k] = -hg[(size_t)i * ld + k];
    }
  }
  cudaFree(dQ); cudaFree(dG); cudaFree(dV);
  return B200NUTS_OK;
}

static int gather_status(b200nuts_engine* e, std::vector<int>& err, std::vector<int>& iter, std::vector<long long>& ng,
                         std::vector<double>& eps, std::vector<int>& rows) {
  const int C = e->C;
  int *d_err, *d_iter, *d_rows;
  long long* d_ng;
  double* d_eps;
  CUDA_TRY(e, cudaSetDevice(e->cfg.device));
  CUDA_TRY(e, cudaMalloc(&d_err, 4 * C));
  CUDA_TRY(e, cudaMalloc(&d_iter, 4 * C));
  CUDA_TRY(e, cudaMalloc(&d_rows, 4 * C));
  CUDA_TRY(e, cudaMalloc(&d_ng, 8 * C));
  CUDA_TRY(e, cudaMalloc(&d_eps, 8 * C));
  gather_status_kernel<<<(C + 255) / 256, 256, 0, e->stream>>>(e->dc.cs, C, d_err, d_iter, d_ng, d_eps, d_rows);
  err.resize(C); iter.resize(C); ng.resize(C); eps.resize(C); rows.resize(C);
  cudaMemcpyAsync(err.data(), d_err, 4 * C, cudaMemcpyDeviceToHost, e->stream);
  cudaMemcpyAsync(iter.data(), d_iter, 4 * C, cudaMemcpyDeviceToHost, e->stream);
  cudaMemcpyAsync(rows.data(), d_rows, 4 * C, cudaMemcpyDeviceToHost, e->stream);
  cudaMemcpyAsync(ng.data(), d_ng, 8 * C, cudaMemcpyDeviceToHost, e->stream);
  cudaMemcpyAsync(eps.data(), d_eps, 8 * C, cudaMemcpyDeviceToHost, e->stream);
  cudaError_t r = cudaStreamSynchronize(e->stream);
  cudaFree(d_err); cudaFree(d_iter); cudaFree(d_rows); cudaFree(d_ng); cudaFree(d_eps);
  if (r != cudaSuccess) { e->err = cudaGetErrorString(r); return B200NUTS_SOFTWARE; }
  return B200NUTS_OK;
}

int b200nuts_get_chain_status(b200nuts_engine* e, int32_t* err, int32_t* iter, int64_t* n_grad) {
  if (!e->chains_ready) { e->err = "chains not initialised"; return B200NUTS_USAGE; }
  std::vector<int> ve, vi, vr;
  std::vector<long long> vg;
  std::vector<double> veps;
  int rc = gather_status(e, ve, vi, vg, veps, vr);
  if (rc) return rc;
  for (int c = 0; c < e->C; c++) {
    if (err) err[c] = ve[c];
    if (iter) iter[c] = vi[c];
    if (n_grad) n_grad[c] = vg[c];
  }
  return B200NUTS_OK;
}

int64_t b200nuts_rows_saved(const b200nuts_engine* ce) {
  auto* e = const_cast<b200nuts_engine*>(ce);
  if (!e->chains_ready) return 0;
  std::vector<int> ve, vi, vr;
  std::vector<long long> vg;
  std::vector<double> veps;
  if (gather_status(e, ve, vi, vg, veps, vr)) return -1;
  int mn = vr.empty() ? 0 : vr[0];
  for (int v : vr) mn = std::min(mn, v);
  return mn;
}

int b200nuts_get_adaptation(b200nuts_engine* e, double* stepsize, double* inv_metric) {
  if (!e->chains_ready) { e->err = "chains not initialised"; return B200NUTS_USAGE; }
  std::vector<int> ve, vi, vr;
  std::vector<long long> vg;
  std::vector<double> veps;
  int rc = gather_status(e, ve, vi, vg, veps, vr);
  if (rc) return rc;
  if (stepsize) std::memcpy(stepsize, veps.data(), sizeof(double) * e->C);
  if (inv_metric) {
    double* d_out;
    const bool dense = e->scfg.dense != 0;
    const long long tot = (long long)e->C * e->D * (dense ? e->D : 1);
    CUDA_TRY(e, cudaMalloc(&d_out, sizeof(double) * tot));
    if (dense) gather_dense_metric_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, e->stream>>>(e->dc.mat, e->dc.mat_stride, e->D, e->C, d_out);
    else gather_metric_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, e->stream>>>(e->dc.slab, e->dc.slab_stride, e->ld, e->D, e->C, d_out);
    cudaMemcpyAsync(inv_metric, d_out, sizeof(double) * tot, cudaMemcpyDeviceToHost, e->stream);
    cudaError_t r = cudaStreamSynchronize(e->stream);
    cudaFree(d_out);
    if (r != cudaSuccess) { e->err = cudaGetErrorString(r); return B200NUTS_SOFTWARE; }
  }
  return B200NUTS_OK;
}

int b200nuts_fetch_draws(b200nuts_engine* e, int64_t chain0, int64_t nchains, int64_t row0, int64_t nrows, int constrained,
                         double* out) {
  if (!e->chains_ready || !e->dc.draws) { e->err = "no draws are stored (store_draws=0 or chains not initialised)"; return B200NUTS_USAGE; }
  if (chain0 < 0 || nchains < 0 || chain0 + nchains > e->C || row0 < 0 || nrows < 0 || row0 + nrows > e->scfg.rows_per_chain) {
    e->err = "fetch_draws: range out of bounds";
    return B200NUTS_USAGE;
  }
  if (nchains == 0 || nrows == 0) return B200NUTS_OK;
  CUDA_TRY(e, cudaSetDevice(e->cfg.device));
  const int kout = constrained ? e->K_out : e->D;
  const size_t row_bytes = sizeof(double) * (7 + kout);
  // chunk over chains so that the staging buffer stays <= ~1 GiB
  int64_t chunk = std::max<int64_t>(1, (int64_t)((1ull << 30) / (row_bytes * (size_t)nrows)));
  chunk = std::min(chunk, nchains);
  // device staging buffer of the (write_array'd) rows: kept in the engine and only ever grown -- a cudaMalloc / cudaFree pair per
  // call was measured at 1 .. 700 ms on the bench boxes (r03j / r03m), more than the copy itself
  if (e->fetch_bytes < row_bytes * nrows * chunk) {
    free_dev(e->d_fetch);
    e->d_fetch = nullptr;
    e->fetch_bytes = 0;
    CUDA_TRY(e, cudaMalloc(&e->d_fetch, row_bytes * nrows * chunk));
    e->fetch_bytes = row_bytes * nrows * chunk;
  }
  double* d_out = e->d_fetch;
  for (int64_t off = 0; off < nchains; off += chunk) {
    const int64_t m = std::min(chunk, nchains - off);
    const long long tot = m * nrows;
    const bool plug_wa = constrained && e->plug && e->plug->write_array;
    fetch_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, e->stream>>>(e->dc.draws, e->dc.draws_stride, e->scfg.ncol, plug_wa ? -1 : e->md.kind, e->D,
                                                                       e->md.N, e->md.K, e->K_out, constrained, chain0 + off, m, row0,
                                                                       nrows, d_out);
    if (plug_wa) {  // sampler columns are in place; the model's own write_array fills the rest (one call per chain: the
                    // rows of a chain are contiguous in the draws buffer, rows of different chains are not)
      for (int64_t c = 0; c < m; c++) {
        const double* q = e->dc.draws + (size_t)(chain0 + off + c) * e->dc.draws_stride + (size_t)row0 * e->scfg.ncol + 7;
        e->plug->write_array(e->plug_model, q, nrows, e->scfg.ncol, d_out + (size_t)c * nrows * (7 + kout) + 7, 7 + kout, (void*)e->stream);
      }
    }
    cudaMemcpyAsync(out + (size_t)off * nrows * (7 + kout), d_out, row_bytes * nrows * m, cudaMemcpyDeviceToHost, e->stream);
    cudaError_t r = cudaStreamSynchronize(e->stream);
    if (r == cudaSuccess) r = cudaGetLastError();
    if (r != cudaSuccess) {
      e->err = std::string("fetch_draws: ") + cudaGetErrorString(r);
      return B200NUTS_SOFTWARE;
    }
  }
  return B200NUTS_OK;
}

int b200nuts_fetch_diagnostics(b200nuts_engine* e, int64_t chain0, int64_t nchains, int64_t row0, int64_t nrows, double* out) {
  if (!e->chains_ready || !e->dc.draws || !e->scfg.diag) { e->err = "no diagnostics are stored (store_diagnostics=0, store_draws=0 or chains not initialised)"; return B200NUTS_USAGE; }
  if (chain0 < 0 || nchains < 0 || chain0 + nchains > e->C || row0 < 0 || nrows < 0 || row0 + nrows > e->scfg.rows_per_chain) {
    e->err = "fetch_diagnostics: range out of bounds";
    return B200NUTS_USAGE;
  }
  if (nchains == 0 || nrows == 0) return B200NUTS_OK;
  CUDA_TRY(e, cudaSetDevice(e->cfg.device));
  // the stored rows are exactly the diagnostic rows: one strided device-to-host copy per chain
  const size_t row_bytes = sizeof(double) * (size_t)e->scfg.ncol;
  for (int64_t c = 0; c < nchains; c++)
    CUDA_TRY(e, cudaMemcpyAsync(out + (size_t)c * nrows * e->scfg.ncol, e->dc.draws + (size_t)(chain0 + c) * e->dc.draws_stride + (size_t)row0 * e->scfg.ncol,
                                row_bytes * nrows, cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(e, cudaStreamSynchronize(e->stream));
  return B200NUTS_OK;
}

int b200nuts_summary_num_stats(void) { return kSummaryStats; }

int b200nuts_summary(b200nuts_engine* e, int64_t row0, int64_t nrows, int constrained, double* out) {
  if (!e->chains_ready || !e->dc.draws) { e->err = "no draws are stored (store_draws=0 or chains not initialised)"; return B200NUTS_USAGE; }
  if (row0 < 0 || nrows < 0 || row0 + nrows > e->scfg.rows_per_chain) { e->err = "summary: range out of bounds"; return B200NUTS_USAGE; }
  CUDA_TRY(e, cudaSetDevice(e->cfg.device));
  if (constrained && e->plug && e->plug->write_array) {
    // plug-in model: materialise its write_array outputs next to the sampler columns (device to device), then summarise
    const int ncol = 7 + e->K_out;
    double* tmp = nullptr;
    CUDA_TRY(e, cudaMalloc(&tmp, sizeof(double) * (size_t)e->C * nrows * ncol));
    const long long tot = (long long)e->C * nrows;
    fetch_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, e->stream>>>(e->dc.draws, e->dc.draws_stride, e->scfg.ncol, -1, e->D, 0, 0, e->K_out, 1,
                                                                       0, e->C, row0, nrows, tmp);
    for (int64_t c = 0; c < e->C; c++)
      e->plug->write_array(e->plug_model, e->dc.draws + (size_t)c * e->dc.draws_stride + (size_t)row0 * e->scfg.ncol + 7, nrows, e->scfg.ncol,
                           tmp + (size_t)c * nrows * ncol + 7, ncol, (void*)e->stream);
    SummaryArgs a;
    a.draws = tmp;
    a.draws_stride = (size_t)nrows * ncol;
    a.ncol = ncol;
    a.kind = MODEL_STD_NORMAL;
    a.D = a.K_out = e->K_out;
    a.N = a.K = 0;
    a.constrained = 0;
    a.C = e->C;
    a.row0 = 0;
    a.n = nrows;
    a.stream = e->stream;
    const int rc = summary_run(a, out, &e->err);
    cudaFree(tmp);
    return rc;
  }
  SummaryArgs a;
  a.draws = e->dc.draws;
  a.draws_stride = e->dc.draws_stride;
  a.ncol = e->scfg.ncol;
  a.kind = e->md.kind;
  a.D = e->D;
  a.N = e->md.N;
  a.K = e->md.K;
  a.K_out = e->K_out;
  a.constrained = constrained;
  a.C = e->C;
  a.row0 = row0;
  a.n = nrows;
  a.stream = e->stream;
  return summary_run(a, out, &e->err);
}

// ---- FP64 FMA peak (roofline denominator of the scalar-FP64-bound kernels)
}  // extern "C"
namespace {
__global__ void __launch_bounds__(256) fma_peak_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {  // 8 independent dependency chains per thread
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}
}  // namespace
extern "C" {
int b200nuts_measure_fp64_fma_peak(int device, double* tflops) {
  int ndev = 0, sms = 0;
  if (!tflops || cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev || cudaSetDevice(device) != cudaSuccess) {
    g_create_error = "measure_fp64_fma_peak: no such CUDA device";
    return B200NUTS_SOFTWARE;
  }
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const int blocks = sms * 8, threads = 256, iters = 1 << 15;
  double* d_out = nullptr;
  cudaEvent_t e0, e1;
  if (cudaMalloc(&d_out, sizeof(double) * (size_t)blocks * threads) != cudaSuccess) { g_create_error = "measure_fp64_fma_peak: cudaMalloc failed"; return B200NUTS_SOFTWARE; }
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 4; rep++) {  // first repetition warms up
    cudaEventRecord(e0);
    fma_peak_kernel<<<blocks, threads>>>(d_out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) break;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms > 0) best = std::max(best, 2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12);
  }
  const cudaError_t err = cudaGetLastError();
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  if (err != cudaSuccess || best <= 0.0) { g_create_error = std::string("measure_fp64_fma_peak: ") + cudaGetErrorString(err); return B200NUTS_SOFTWARE; }
  *tflops = best;
  return B200NUTS_OK;
}

// ---- plug-in registry (include/b200nuts_model.h) -----------------------------------------------------------------------
int b200nuts_register_model(const b200nuts_model_vtable* vt) {
  if (!vt || vt->abi < 1 || vt->abi > B200NUTS_MODEL_ABI || !vt->name || !vt->create || !vt->log_prob_grad) {
    g_create_error = "register_model: bad vtable (abi, name, create and log_prob_grad are required)";
    return -B200NUTS_USAGE;
  }
  std::lock_guard<std::mutex> lk(g_plugin_mutex);
  for (size_t i = 0; i < g_plugins.size(); i++)
    if (std::strcmp(g_plugins[i]->name, vt->name) == 0) return B200NUTS_MODEL_PLUGIN_BASE + (int)i;
  for (const char* b : kBuiltinNames)
    if (std::strcmp(b, vt->name) == 0) { g_create_error = std::string("register_model: '") + vt->name + "' is a built-in model name"; return -B200NUTS_USAGE; }
  g_plugins.push_back(vt);
  return B200NUTS_MODEL_PLUGIN_BASE + (int)g_plugins.size() - 1;
}

int b200nuts_load_model_library(const char* path) {
  void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!h) {
    const char* m = dlerror();
    g_create_error = std::string("load_model_library: ") + (m ? m : path);
    return -B200NUTS_NOINPUT;
  }
  auto entry = (b200nuts_model_entry_fn)dlsym(h, "b200nuts_model_entry");
  if (!entry) { g_create_error = std::string("load_model_library: ") + path + " does not export b200nuts_model_entry"; return -B200NUTS_USAGE; }
  int first = -1;
  for (int i = 0;; i++) {
    const b200nuts_model_vtable* vt = entry(i);
    if (!vt) break;
    const int id = b200nuts_register_model(vt);
    if (id < 0) return id;
    if (first < 0) first = id;
  }
  if (first < 0) { g_create_error = std::string("load_model_library: ") + path + " exports no models"; return -B200NUTS_USAGE; }
  return first;  // the library stays loaded for the life of the process (its vtables are referenced by the registry)
}

int b200nuts_find_model(const char* name) {
  for (int i = 0; i < 6; i++)
    if (std::strcmp(kBuiltinNames[i], name) == 0) return i;
  std::lock_guard<std::mutex> lk(g_plugin_mutex);
  for (size_t i = 0; i < g_plugins.size(); i++)
    if (std::strcmp(g_plugins[i]->name, name) == 0) return B200NUTS_MODEL_PLUGIN_BASE + (int)i;
  return -1;
}

// constrained_param_names (include_tparams = true) of the built-in family: same names as the host model objects
// (cmdstan_b200/host/model.hpp, cmdstan_b200.param_names)
int b200nuts_output_name(b200nuts_engine* e, int64_t k, char* buf, int buflen) {
  int rc = finalize_data(e);
  if (rc) return rc;
  if (k < 0 || k >= e->K_out || !buf || buflen < 2) { e->err = "output_name: index out of range"; return B200NUTS_USAGE; }
  if (e->plug) {
    if (e->plug->output_name) return e->plug->output_name(e->plug_model, k, buf, buflen);
    std::snprintf(buf, (size_t)buflen, "v%lld", (long long)k + 1);
    return B200NUTS_OK;
  }
  const long long N = e->md.N, K = e->md.K, i = k;
  auto vec = [&](const char* n, long long j) { std::snprintf(buf, (size_t)buflen, "%s.%lld", n, j + 1); };
  auto sc = [&](const char* n) { std::snprintf(buf, (size_t)buflen, "%s", n); };
  switch (e->md.kind) {
    case MODEL_STD_NORMAL: vec("x", i); break;
    case MODEL_BERNOULLI: sc("theta"); break;
    case MODEL_EIGHT_SCHOOLS:
      if (i == 0) sc("mu"); else if (i == 1) sc("tau"); else if (i < 2 + N) vec("theta_tilde", i - 2); else vec("theta", i - 2 - N);
      break;
    case MODEL_LOGISTIC:
    case MODEL_POISSON:
      if (i == 0) sc("alpha"); else vec("beta", i - 1);
      break;
    case MODEL_IRT2PL:
      if (i == 0) sc("sigma_theta"); else if (i < 1 + N) vec("theta", i - 1); else if (i == 1 + N) sc("sigma_a");
      else if (i < 2 + N + K) vec("a", i - 2 - N); else if (i == 2 + N + K) sc("mu_b"); else if (i == 3 + N + K) sc("sigma_b");
      else vec("b", i - 4 - N - K);
      break;
    default: vec("v", i);
  }
  return B200NUTS_OK;
}

int b200nuts_transform_inits(b200nuts_engine* e, const double* x, int64_t n, double* q) {
  int rc = finalize_data(e);
  if (rc) return rc;
  const int64_t D = e->D, K = e->K_out;
  if (e->plug) {
    if (e->plug->abi < 2 || !e->plug->transform_inits) {
      e->err = std::string(e->plug->name) + ": the model has no transform_inits (user-specified initial values are not available; use init=<radius>)";
      return B200NUTS_USAGE;
    }
    char msg[512] = {0};
    for (int64_t c = 0; c < n; c++) {
      rc = e->plug->transform_inits(e->plug_model, x + c * K, q + c * D, msg, (int)sizeof msg);
      if (rc) {
        e->err = msg[0] ? msg : "transform_inits failed";
        return B200NUTS_CONFIG;
      }
    }
    return B200NUTS_OK;
  }
  // built-in models: the parameters are the first D output columns, element-wise transforms (M/prim/constraint/lb_free.hpp, lub_free.hpp)
  const int I = e->md.N, J = e->md.K;
  auto kind_of = [&](int64_t i) -> int {  // 0 identity, 1 lower bound 0, 2 unit interval
    switch (e->md.kind) {
      case MODEL_BERNOULLI: return 2;
      case MODEL_EIGHT_SCHOOLS: return i == 1 ? 1 : 0;
      case MODEL_IRT2PL: return (i == 0 || i == 1 + I || (i >= 2 + I && i < 2 + I + J) || i == 3 + I + J) ? 1 : 0;
      default: return 0;
    }
  };
  for (int64_t c = 0; c < n; c++)
    for (int64_t i = 0; i < D; i++) {
      const double v = x[c * K + i];
      const int kd = kind_of(i);
      double u;
      if (v != v) {
        u = b200nuts_init_not_given(kd);
      } else if (kd == 1) {
        if (!(v >= 0)) { e->err = "lb_free: Bounded variable is " + std::to_string(v) + ", but must be greater than or equal to 0"; return B200NUTS_CONFIG; }
        u = std::log(v);
      } else if (kd == 2) {
        if (!(v >= 0 && v <= 1)) { e->err = "lub_free: Bounded variable is " + std::to_string(v) + ", but must be in the interval [0, 1]"; return B200NUTS_CONFIG; }
        u = std::log(v / (1 - v));
      } else {
        u = v;
      }
      q[c * D + i] = u;
    }
  return B200NUTS_OK;
}

int b200nuts_summarize_draws(const double* draws, int64_t chains, int64_t rows, int64_t ncols, int device, double* out, char* errbuf,
                             int errlen) {
  std::string err;
  int rc = B200NUTS_OK;
  double* d = nullptr;
  cudaStream_t st = nullptr;
  auto fail = [&](const std::string& m, int code) {
    err = m;
    rc = code;
  };
  int ndev = 0;
  if (chains < 1 || rows < 2 || ncols < 1) fail("summarize_draws: chains >= 1, rows >= 2, ncols >= 1 required", B200NUTS_USAGE);
  else if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) fail("no CUDA device available (libb200nuts has no CPU path)", B200NUTS_SOFTWARE);
  else if (device < 0 || device >= ndev) fail("device ordinal out of range", B200NUTS_USAGE);
  else if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) fail("cudaSetDevice / cudaStreamCreate failed", B200NUTS_SOFTWARE);
  else {
    const size_t bytes = sizeof(double) * (size_t)chains * rows * ncols;
    if (cudaMalloc(&d, bytes) != cudaSuccess || cudaMemcpyAsync(d, draws, bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) {
      fail("summarize_draws: cannot upload the draws", B200NUTS_SOFTWARE);
    } else {
      SummaryArgs a;
      a.draws = d;
      a.draws_stride = (size_t)rows * ncols;
      a.ncol = (int)ncols;
      a.kind = MODEL_STD_NORMAL;
      a.D = (int)ncols - 7;
      a.N = a.K = 0;
      a.K_out = a.D;
      a.constrained = 0;
      a.C = chains;
      a.row0 = 0;
      a.n = rows;
      a.stream = st;
      rc = summary_run(a, out, &err);
    }
  }
  if (d) cudaFree(d);
  if (st) cudaStreamDestroy(st);
  if (rc && errbuf && errlen > 0) {
    std::strncpy(errbuf, err.c_str(), (size_t)errlen - 1);
    errbuf[errlen - 1] = 0;
  }
  return rc;
}

}  // extern "C"
