// Hierarchical IRT 2PL log-density + gradient for the lock-step engine (config c4), one CTA per requesting chain.
//
//   parameters (declaration order = unconstrained layout, positive ones as logs with Jacobian +u, lb_constrain.hpp:97):
//     sigma_theta, theta[I], sigma_a, a[J], mu_b, sigma_b, b[J]                              D = 4 + I + 2J
//   model: sigma_* ~ cauchy(0,2) (cauchy_lpdf.hpp:80-86); theta ~ normal(0, sigma_theta), b ~ normal(mu_b, sigma_b)
//          (normal_lpdf.hpp:83-89, keeps -N log sigma); a ~ lognormal(0, sigma_a) (lognormal_lpdf.hpp:72-80, keeps
//          N * NEG_LOG_SQRT_TWO_PI even under propto); mu_b ~ normal(0,5);
//          y[i,j] ~ bernoulli_logit(a_j (theta_i - b_j)) (bernoulli_logit_lpmf.hpp:73-89, cutoff +-20).
//
// Work distribution: the CTA's 256 threads form 32 groups of 8 lanes; a group takes persons i = g, g+32, ...; inside a
// group lane s owns items j = s, s+8, ... (IPL = ceil(J/8) items per lane: J = 100 -> 13, 96 % of the lane slots busy),
// whose a/b gradients accumulate in registers over all the group's persons; the theta_i gradient is a 3-step shuffle
// reduction over the group.  The per-response link is the branch-free one of fastmath.cuh (the kernel is bound by the
// FP64 pipe: ~51 FP64 instructions per response, 2 I J of them transcendental-class).  y is shared by all chains (int8,
// L2-resident); per chain the kernel reads q (D doubles) and writes g (D doubles).  All sums run in a fixed order
// (shuffle trees, then warp order) => deterministic and independent of the batch.
#include <cstdlib>

#include "irt_body.cuh"

namespace b200 {

int irt_max_items() { return IRT_G * IRT_MAX_IPL; }

// OCC = 1: item parameters in registers, one CTA per SM (254 registers).  OCC = 2: exp(a), b in shared memory and a 128-register
// budget, two CTAs (16 warps) per SM: the r01c capture showed the one-CTA form latency-bound (FP64 pipe 47 % busy, 45 % of
// the samples in fixed-latency dependency stalls at 2 warps per sub-partition).
template <int IPL, int OCC>
__global__ void __launch_bounds__(IRT_NT, OCC) irt_grad_kernel(const signed char* __restrict__ y, int I, int J, int D,
                                                             const double* __restrict__ Q, double* __restrict__ G,
                                                             double* __restrict__ zV, const int* __restrict__ active,
                                                             const int* __restrict__ list, const int* __restrict__ n_list, int ld) {
  int c = blockIdx.x;
  if (list) {
    if (c >= *n_list) return;
    c = list[c];
  } else if (active && !active[c]) {
    return;
  }
  irt_chain_body<IPL, OCC>(y, I, J, D, Q + (size_t)c * ld, G + (size_t)c * ld, zV + c);
}

int launch_irt_grad(const ModelData& m, const GradBatch& b, cudaStream_t stream) {
  const int ipl = (m.K + IRT_G - 1) / IRT_G;
  // B200NUTS_IRT_VARIANT: 1 (default) one CTA per SM; 2 two CTAs per SM; "auto" (EXPERIMENTAL: needs the two forms to be
  // bit-identical, tests/test_gpu_parity.py with B200_TEST_EXPERIMENTAL=1) picks 2 when the launch is known to be dense
  // (faster at full batch, profiles r01d) and 1 when only stragglers are left (lower single-CTA latency, profiles r01e).
  // Read per launch (a handful of ns next to a > 100 us kernel): tests switch it.
  const char* var = getenv("B200NUTS_IRT_VARIANT");
  int occ = var ? atoi(var) : 1;
  if (var && var[0] == 'a') occ = (b.hint_active < 0 ? b.C : b.hint_active) >= 296 ? 2 : 1;
#define IRT_LAUNCH(IPL)                                                                                                          \
  {                                                                                                                              \
    auto k1 = irt_grad_kernel<IPL, 1>;                                                                                           \
    auto k2 = irt_grad_kernel<IPL, 2>;                                                                                           \
    if (occ == 2) IRT_KERNEL_LAUNCH(k2, b.C, stream, m.y8, m.N, m.K, m.D, b.Q, b.G, b.zV, b.active, b.list, b.n_list, b.ld);     \
    else IRT_KERNEL_LAUNCH(k1, b.C, stream, m.y8, m.N, m.K, m.D, b.Q, b.G, b.zV, b.active, b.list, b.n_list, b.ld);              \
  }
  if (ipl <= 2) IRT_LAUNCH(2)
  else if (ipl <= 4) IRT_LAUNCH(4)
  else if (ipl <= 7) IRT_LAUNCH(7)
  else if (ipl <= 10) IRT_LAUNCH(10)
  else if (ipl <= 13) IRT_LAUNCH(13)
  else IRT_LAUNCH(16)
#undef IRT_LAUNCH
  return 1;
}

}  // namespace b200
