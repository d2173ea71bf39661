// TEST-ONLY: the chain kernels of the engine (cmdstan_b200/csrc/chains.cuh: fused_chain_kernel with the state in shared or
// global memory, advance_kernel of the lock-step engine) executed on host threads with REAL 32-lane warps by the SIMT stand-in
// under tests/host_emul/simt.  tests/host_emul/emul.cpp drives the same state machine with a one-lane warp (bit-identical to
// the oracle); this rig exercises what that cannot: lane-strided vector loops, the shuffle reductions, the owner-lane rule
// (ThreadSanitizer build: a lane that touches another lane's element without a __syncwarp is a data race), the shared-memory
// staging of the chain state, the compacted request list.  Built like dense.cu (every optional feature compiled in).
#define B200_SIMT_EMUL 1
#define B200_WITH_DENSE 1
#include <cuda_runtime.h>  // tests/host_emul/simt/cuda_runtime.h

#include "../../cmdstan_b200/csrc/chains.cuh"
#include "../../cmdstan_b200/csrc/host_seed.h"

using namespace b200;

namespace {
// engine.cu's grad_small_kernel (gradient stage of the lock-step engine for the small models)
void grad_small_body(ModelData md, const double* Q, double* G, double* zV, const int* active, int C, int ld) {
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
}  // namespace

// mode 0: fused kernel, state in global memory; 1: fused kernel, state staged in shared memory; 2: lock-step (advance kernel +
// gradient kernel per leapfrog).  q0: [C][D] or null; inv_metric0: [C][D] (diag) / [C][D][D] (dense) or null.
// draws: [C][rows][ncol]; adapt_out: [C][1 + D (+ D D)]; n_grad, err: [C].  Returns the number of launches.
extern "C" long long simt_run_chains(int mode, int C, int kind, int D, int N, const double* yr, const double* sigma, int sum_y, unsigned seed,
                                     unsigned chain_id, int num_warmup, int num_samples, int num_thin, int save_warmup, double stepsize,
                                     double jitter, int max_depth, double delta, double gamma, double kappa, double t0, unsigned init_buffer,
                                     unsigned term_buffer, unsigned window, double init_radius, int adapt_engaged, const double* q0,
                                     const double* inv_metric0, double* draws, double* adapt_out, long long* n_grad, int* err, int unit_e,
                                     int dense_e, int diag, int static_hmc, double int_time) {
  ModelData m{};
  m.kind = kind; m.D = D; m.N = N; m.yr = yr; m.sigma = sigma; m.sum_y = sum_y;
  const int ld = (D + 3) & ~3;
  SamplerCfg cfg{};
  cfg.D = D; cfg.ld = ld; cfg.max_depth = max_depth;
  cfg.num_warmup = num_warmup; cfg.num_samples = num_samples; cfg.num_thin = num_thin; cfg.save_warmup = save_warmup;
  cfg.adapt_engaged = adapt_engaged; cfg.user_inits = q0 != nullptr;
  cfg.stepsize = stepsize; cfg.stepsize_jitter = (jitter > 0 && jitter < 1) ? jitter : 0.0;
  cfg.delta = (delta > 0 && delta < 1) ? delta : 0.5; cfg.gamma = gamma > 0 ? gamma : 0.05;
  cfg.kappa = kappa > 0 ? kappa : 0.75; cfg.t0 = t0 > 0 ? t0 : 10; cfg.init_radius = init_radius;
  cfg.dense = dense_e; cfg.diag = diag; cfg.static_hmc = static_hmc; cfg.int_time = int_time;
  resolve_window(&cfg, (adapt_engaged && !unit_e) ? num_warmup : 0, init_buffer, term_buffer, window);
  cfg.ncol = 7 + D * (diag ? 3 : 1);
  cfg.rows_per_chain = (save_warmup ? (num_warmup + num_thin - 1) / num_thin : 0) + (num_samples + num_thin - 1) / num_thin;

  DeviceChains d{};
  d.slab_stride = (size_t)num_vec_slots(max_depth, dense_e, diag) * ld;
  d.mat_stride = dense_e ? 3 * (size_t)D * D : 0;
  d.draws_stride = (size_t)cfg.rows_per_chain * cfg.ncol;
  std::vector<ChainS> cs(C);
  std::vector<double> q((size_t)C * ld, 0.0), g((size_t)C * ld, 0.0), zV(C, 0.0), slab((size_t)C * d.slab_stride, 0.0), mat((size_t)C * d.mat_stride + 1, 0.0),
      uq;
  for (int c = 0; c < C; c++) {
    std::memset(&cs[c], 0, sizeof(ChainS));
    chain_init_scalars(&cs[c], cfg);
    seed_stream(cs[c].rng, seed, chain_id + (unsigned)c);
    cs[c].iter_end = num_warmup + num_samples;
    if (dense_e) {  // column-major copy of the (symmetric) inverse metric, identity by default (engine.cu: init_dense_metric_kernel)
      for (int i = 0; i < D; i++)
        for (int j = 0; j < D; j++)
          mat[(size_t)c * d.mat_stride + (size_t)j * D + i] = inv_metric0 ? inv_metric0[((size_t)c * D + i) * D + j] : (i == j ? 1.0 : 0.0);
    } else {
      for (int i = 0; i < D; i++) slab[(size_t)c * d.slab_stride + (size_t)V_IM * ld + i] = inv_metric0 ? inv_metric0[(size_t)c * D + i] : 1.0;
    }
  }
  if (q0) {
    uq.assign((size_t)C * ld, 0.0);
    for (int c = 0; c < C; c++) std::memcpy(&uq[(size_t)c * ld], q0 + (size_t)c * D, sizeof(double) * D);
  }
  d.cs = cs.data(); d.q = q.data(); d.g = g.data(); d.zV = zV.data(); d.slab = slab.data(); d.draws = draws;
  d.user_q0 = q0 ? uq.data() : nullptr;
  d.mat = dense_e ? mat.data() : nullptr;

  long long launches = 0;
  if (mode == 0) {
    simt::launch((C + kChainsPerBlock - 1) / kChainsPerBlock, 32 * kChainsPerBlock, 0, [&] { fused_chain_kernel<false, 2>(cfg, m, d, C, kChainsPerBlock, 0); });
    launches = 1;
  } else if (mode == 1) {
    const int chain_words = kChainWords + 2 * ld + 2 + (int)d.slab_stride;
    int wpb = C < 3 ? C : 3;  // engine.cu spreads a per-SM wave over blocks of <= 7 warps; any split must give the same chains
    simt::launch((C + wpb - 1) / wpb, 32 * wpb, (size_t)wpb * chain_words * 8, [&] { fused_chain_kernel<true, 2>(cfg, m, d, C, wpb, chain_words); });
    launches = 1;
  } else {
    std::vector<int> active(C, 0), list(C, 0);
    for (;;) {
      int n_active = 0;
      simt::launch((C + kChainsPerBlock - 1) / kChainsPerBlock, 32 * kChainsPerBlock, 0, [&] { advance_kernel<2>(cfg, d, C, active.data(), list.data(), &n_active); });
      launches++;
      if (n_active == 0) break;
      int cnt = 0;
      for (int c = 0; c < C; c++) cnt += active[c];
      if (cnt != n_active) simt::die("advance_kernel: request list and active flags disagree");
      simt::launch((C + 3) / 4, 128, 0, [&] { grad_small_body(m, q.data(), g.data(), zV.data(), active.data(), C, ld); });
      launches++;
    }
  }
  const size_t ao = 1 + (size_t)D + (dense_e ? (size_t)D * D : 0);
  for (int c = 0; c < C; c++) {
    adapt_out[c * ao] = cs[c].eps_nom;
    if (dense_e) {
      for (int i = 0; i < D; i++)
        for (int j = 0; j < D; j++) adapt_out[c * ao + 1 + (size_t)i * D + j] = mat[(size_t)c * d.mat_stride + (size_t)j * D + i];
    } else {
      for (int i = 0; i < D; i++) adapt_out[c * ao + 1 + i] = slab[(size_t)c * d.slab_stride + (size_t)V_IM * ld + i];
    }
    n_grad[c] = cs[c].n_grad;
    err[c] = cs[c].err;
  }
  return launches;
}

#ifdef CHAINS_SIMT_MAIN
// stand-alone form for ThreadSanitizer (g++ -fsanitize=thread -DCHAINS_SIMT_MAIN): eight schools / std-normal chains, a checksum
int main(int argc, char** argv) {
  if (argc < 8) {
    std::fprintf(stderr, "usage: %s mode kind D dense diag static num_iter\n", argv[0]);
    return 2;
  }
  const int mode = std::atoi(argv[1]), kind = std::atoi(argv[2]), D = std::atoi(argv[3]), dense = std::atoi(argv[4]), diag = std::atoi(argv[5]),
            stat = std::atoi(argv[6]), n = std::atoi(argv[7]);
  const double y[8] = {28, 8, -3, 7, -1, 1, 18, 12}, sigma[8] = {15, 10, 16, 11, 9, 11, 10, 18};
  const int C = 3, rows = 2 * n, ncol = 7 + D * (diag ? 3 : 1);
  std::vector<double> draws((size_t)C * rows * ncol, 0.0), adapt((size_t)C * (1 + D + (dense ? D * D : 0)), 0.0);
  std::vector<long long> ng(C);
  std::vector<int> err(C);
  simt_run_chains(mode, C, kind, D, kind == 2 ? 8 : 0, y, sigma, 0, 11u, 1u, n, n, 1, 1, 1.0, 0.0, 5, 0.8, 0.05, 0.75, 10.0, 75u, 50u, 25u, 2.0, 1, nullptr, nullptr,
                  draws.data(), adapt.data(), ng.data(), err.data(), 0, dense, diag, stat, 1.5);
  double sum = 0;
  for (double v : draws) sum += v;
  std::printf("draws checksum %.17g  err %d %d %d\n", sum, err[0], err[1], err[2]);
  return 0;
}
#endif
