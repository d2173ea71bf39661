// TEST-ONLY: the IRT 2PL kernels (cmdstan_b200/csrc/irt.cu: the throughput gradient kernel in its one- and two-CTA-per-SM forms,
// and the latency kernel irt_solo_kernel in its one-CTA-per-chain form -- the cluster form needs distributed shared memory, which
// the rig does not model; both forms share the block arithmetic, the CTA-cooperative advance step and the hand-over logic)
// executed on host threads by the SIMT stand-in under tests/host_emul/simt, through the product's launchers.
// tests/test_irt_kernels_simt.py: chains finished by the latency kernel must equal chains finished by the lock-step engine bit
// for bit.
#define B200_SIMT_EMUL 1
#define B200_FASTMATH_HOST 1
#include <cuda_runtime.h>  // tests/host_emul/simt/cuda_runtime.h

#include "../../cmdstan_b200/csrc/irt.cu"
#include "../../cmdstan_b200/csrc/advance_cta.cu"
#include "../../cmdstan_b200/csrc/host_seed.h"

using namespace b200;

// form 1 / 2: irt_grad_kernel with one / two CTAs per SM.  Q, G: [C][ld]; zV: [C].
extern "C" int irt_simt_grad(int form, int I, int J, int C, int ld, const signed char* y, const double* Q, const int* active, const int* list, int n_list,
                             double* G, double* zV) {
  ModelData m{};
  m.kind = MODEL_IRT2PL;
  m.N = I; m.K = J; m.D = 4 + I + 2 * J; m.y8 = y;
  std::vector<unsigned short> ypk_((size_t)I * 8);
  irt_pack_responses(y, I, J, ypk_.data());
  m.ypk = ypk_.data();
  GradBatch b{};
  b.Q = Q; b.G = G; b.zV = zV; b.active = active; b.C = C; b.ld = ld;
  int nl = n_list;
  if (list) { b.list = list; b.n_list = &nl; }
  setenv("B200NUTS_IRT_VARIANT", form == 2 ? "2" : "1", 1);
  return launch_irt_grad(m, b, nullptr);
}

static int g_advance_cta = 0;
extern "C" void irt_simt_set_advance_cta(int on) { g_advance_cta = on; }

// C chains of the IRT model through the lock-step engine (advance kernel + irt_grad_kernel per leapfrog); after
// `tail_after` lock-steps (< 0: never) the chains still requesting a gradient are handed to irt_solo_kernel (one CTA per chain),
// which runs each of them to the end in one launch.  draws: [C][rows][7 + D].  Returns the number of launches.
extern "C" long long irt_simt_run(int C, int I, int J, const signed char* y, unsigned seed, int num_warmup, int num_samples, int max_depth,
                                  long long tail_after, double* draws, long long* n_grad, int* err) {
  ModelData m{};
  m.kind = MODEL_IRT2PL;
  m.N = I; m.K = J; m.D = 4 + I + 2 * J; m.y8 = y;
  std::vector<unsigned short> ypk_((size_t)I * 8);
  irt_pack_responses(y, I, J, ypk_.data());
  m.ypk = ypk_.data();
  const int D = m.D, ld = (D + 3) & ~3;
  SamplerCfg cfg{};
  cfg.D = D; cfg.ld = ld; cfg.max_depth = max_depth;
  cfg.num_warmup = num_warmup; cfg.num_samples = num_samples; cfg.num_thin = 1; cfg.save_warmup = 1;
  cfg.adapt_engaged = 1; cfg.stepsize = 1.0; cfg.delta = 0.8; cfg.gamma = 0.05; cfg.kappa = 0.75; cfg.t0 = 10; cfg.init_radius = 2.0;
  resolve_window(&cfg, num_warmup, 75, 50, 25);
  cfg.ncol = 7 + D;
  cfg.rows_per_chain = num_warmup + num_samples;
  DeviceChains d{};
  d.slab_stride = (size_t)num_vec_slots(max_depth, 0, 0) * ld;
  d.draws_stride = (size_t)cfg.rows_per_chain * cfg.ncol;
  std::vector<ChainS> cs(C);
  std::vector<double> q((size_t)C * ld, 0.0), g((size_t)C * ld, 0.0), zV(C, 0.0), slab((size_t)C * d.slab_stride, 0.0);
  for (int c = 0; c < C; c++) {
    std::memset(&cs[c], 0, sizeof(ChainS));
    chain_init_scalars(&cs[c], cfg);
    seed_stream(cs[c].rng, seed, 1u + (unsigned)c);
    cs[c].iter_end = num_warmup + num_samples;
    for (int i = 0; i < D; i++) slab[(size_t)c * d.slab_stride + (size_t)V_IM * ld + i] = 1.0;
  }
  d.cs = cs.data(); d.q = q.data(); d.g = g.data(); d.zV = zV.data(); d.slab = slab.data(); d.draws = draws;
  std::vector<int> active(C, 0), list(C, 0);
  GradBatch b{};
  b.Q = q.data(); b.G = g.data(); b.zV = zV.data(); b.active = active.data(); b.C = C; b.ld = ld; b.list = list.data();
  long long launches = 0;
  for (long long step = 0;; step++) {
    int n_active = 0;
    if (g_advance_cta) launch_advance_cta(cfg, d, C, active.data(), list.data(), &n_active, nullptr);   // one CTA per chain (advance_cta.cu)
    else simt::launch((C + kChainsPerBlock - 1) / kChainsPerBlock, 32 * kChainsPerBlock, 0, [&] { advance_kernel<1>(cfg, d, C, active.data(), list.data(), &n_active); });
    launches++;
    if (n_active == 0) break;
    b.n_list = &n_active;
    if (tail_after >= 0 && step >= tail_after) {
      launch_irt_solo(cfg, d, m, list.data(), &n_active, n_active, 1, nullptr);
      launches++;
      continue;  // one more advance launch finds every chain idle
    }
    launches += launch_irt_grad(m, b, nullptr);
  }
  for (int c = 0; c < C; c++) {
    n_grad[c] = cs[c].n_grad;
    err[c] = cs[c].err;
  }
  return launches;
}

#ifdef IRT_SIMT_MAIN
// stand-alone form for ThreadSanitizer (g++ -fsanitize=thread -DIRT_SIMT_MAIN): synthetic responses, the same chains through the
// lock-step engine and through the straggler kernel, checksums compared
#include <random>
int main(int argc, char** argv) {
  if (argc < 7) {
    std::fprintf(stderr, "usage: %s C I J num_warmup num_samples tail_after\n", argv[0]);
    return 2;
  }
  const int C = std::atoi(argv[1]), I = std::atoi(argv[2]), J = std::atoi(argv[3]), nw = std::atoi(argv[4]), ns = std::atoi(argv[5]);
  const long long tail_after = std::atoll(argv[6]);
  std::mt19937_64 gen(3);
  std::vector<signed char> y((size_t)I * J);
  for (auto& v : y) v = (signed char)(gen() & 1);
  const int D = 4 + I + 2 * J;
  std::vector<double> d0((size_t)C * (nw + ns) * (7 + D), 0.0), d1 = d0;
  std::vector<long long> ng(C);
  std::vector<int> err(C);
  irt_simt_run(C, I, J, y.data(), 5u, nw, ns, 3, -1, d0.data(), ng.data(), err.data());
  irt_simt_run(C, I, J, y.data(), 5u, nw, ns, 3, tail_after, d1.data(), ng.data(), err.data());
  std::printf("lock-step and straggler kernel %s, err %d\n", d0 == d1 ? "agree" : "DIFFER", err[0]);
  return d0 == d1 ? 0 : 1;
}
#endif
