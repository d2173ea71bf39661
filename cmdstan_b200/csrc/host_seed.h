// Host-side helpers shared by the engine: MIXMAX stream seeding (the big skip is done once per chain on the host and
// the 19-word state is uploaded) and the resolution of the adaptation-window schedule.
//   seeding   B/random/mixmax.hpp:220-301 (seed_uniquestream -> apply_bigskip, fmodmulM61)
//   windows   S/mcmc/windowed_adaptation.hpp:23-79 (ctor, restart, set_window_params)
#pragma once
#include <climits>
#include <cstdint>

#include "nuts.cuh"

namespace b200 {

static const unsigned long long kMixmaxSkip[128][17] = {
#include "mixmax_skip_n17.inc"
};

inline unsigned long long mm_fmodmul(unsigned long long cum, unsigned long long s, unsigned long long a) {
  const unsigned long long MASK32 = 0xFFFFFFFFULL;
  unsigned long long o = s * a, ph = s >> 32, pl = s & MASK32, ah = a >> 32, al = a & MASK32;
  o = (o & kM61) + ((ph * ah) << 3) + ((ah * pl + al * ph + ((al * pl) >> 32)) >> 29);
  o += cum;
  o = (o & kM61) + (o >> 61);
  return o;
}

// state of create_rng(seed, chain): unit vector e0 skipped ahead by A^(sum over set ID bits b of 2^(512+b)),
// ID = {streamID = chain, runID = seed, machineID = 1, clusterID = 0} (low word first)
inline void seed_stream(unsigned long long* st, uint32_t seed, uint32_t chain) {
  const uint32_t id[4] = {chain, seed, 1u, 0u};
  unsigned long long Y[17] = {0}, cum[17], sumtot = 1;
  Y[0] = 1;
  for (int k = 0; k < 4; k++) {
    uint32_t v = id[k];
    for (int r = 0; v; v >>= 1, r++) {
      if (!(v & 1)) continue;
      const unsigned long long* row = kMixmaxSkip[r + 32 * k];
      for (int i = 0; i < 17; i++) cum[i] = 0;
      for (int j = 0; j < 17; j++) {
        for (int i = 0; i < 17; i++) cum[i] = mm_fmodmul(cum[i], row[j], Y[i]);
        sumtot = mm_iterate(Y, sumtot);
      }
      sumtot = 0;
      for (int i = 0; i < 17; i++) {
        Y[i] = cum[i];
        sumtot = mm_mod(sumtot + cum[i]);
      }
    }
  }
  sumtot = 0;
  for (int i = 0; i < 17; i++) {
    st[i] = Y[i];
    sumtot = mm_mod(sumtot + Y[i]);
  }
  st[17] = sumtot;
  st[18] = 1;
}

// windowed_adaptation: constructor leaves (0,0,0,0) with next = UINT_MAX; set_window_params either keeps that
// (num_warmup < 20), switches to the 15/75/10 split *without* restart() (so the window end is never reached --
// reference behaviour, windowed_adaptation.hpp:42-67), or installs the given schedule and restarts.
inline void resolve_window(SamplerCfg* cfg, unsigned num_warmup, unsigned init_buffer, unsigned term_buffer, unsigned base_window) {
  cfg->w_num_warmup = cfg->w_init_buffer = cfg->w_term_buffer = cfg->w_base_window = 0;
  cfg->w_size0 = 0;
  cfg->w_next0 = UINT_MAX;
  if (num_warmup < 20) return;
  if (init_buffer + base_window + term_buffer > num_warmup) {
    cfg->w_num_warmup = num_warmup;
    cfg->w_init_buffer = (unsigned)(0.15 * num_warmup);
    cfg->w_term_buffer = (unsigned)(0.10 * num_warmup);
    cfg->w_base_window = num_warmup - (cfg->w_init_buffer + cfg->w_term_buffer);
    return;
  }
  cfg->w_num_warmup = num_warmup;
  cfg->w_init_buffer = init_buffer;
  cfg->w_term_buffer = term_buffer;
  cfg->w_base_window = base_window;
  cfg->w_size0 = base_window;
  cfg->w_next0 = init_buffer + base_window - 1;
}

}  // namespace b200
