// TEST-ONLY stand-in for <cuda_runtime.h>: runs __global__ kernels of cmdstan_b200/csrc on HOST THREADS so that the CPU
// test-suite can execute the data-heavy gradient kernels (DMMA fragments, TMA + mbarrier pipelines, warp shuffles, the
// last-arriver refill protocol) without a GPU, and so that ThreadSanitizer can look for shared-memory races in them.
//
// Model: one std::thread per CUDA thread; the CTAs of a grid run one after another (so `__shared__` can be `static`).
//   __syncthreads / __syncwarp   -> mutex + condition-variable barriers (block / warp)
//   __shfl_sync, __shfl_xor_sync -> post to a per-warp slot, barrier, read the source lane's slot (two slots alternate, one
//                                   barrier per collective)
//   mma.sync.m8n8k4.f64          -> the PTX ISA fragment layout: a = A[lane>>2][lane&3], b = B[lane&3][lane>>2],
//                                   d0,d1 = D[lane>>2][2 (lane&3) + {0,1}];  D = A B + C as an FMA chain over k
//   mbarrier                     -> pending-arrival count + transaction-byte count + phase bit per barrier object
//   cp.async.bulk (1-D TMA)      -> memcpy by the issuing thread, then complete_tx on the barrier
// Every blocking wait has a time-out that aborts with a message: a protocol deadlock fails the test instead of hanging it.
// The kernel sources opt in with B200_SIMT_EMUL (a handful of #ifdefs around inline PTX and <<<>>>).  Never shipped.
#pragma once
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <dlfcn.h>
#include <execinfo.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __restrict__
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static

using std::isfinite;
using std::isnan;

typedef void* cudaStream_t;
typedef int cudaError_t;
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <class F>
inline int cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return 0; }

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct double2 { double x, y; };
inline double2 make_double2(double x, double y) { return double2{x, y}; }

namespace simt {

struct Dim3 { unsigned x = 1, y = 1, z = 1; };

[[noreturn]] inline void die(const char* what) {
  std::fprintf(stderr, "simt: %s\n", what);
  std::fflush(stderr);
  std::_Exit(86);
}

inline int wait_seconds() {  // SIMT_WAIT_SECONDS: shorter when hunting a deadlock
  static const int v = std::getenv("SIMT_WAIT_SECONDS") ? std::atoi(std::getenv("SIMT_WAIT_SECONDS")) : 120;
  return v;
}

// sense-reversing barrier on two atomics; waiters spin with sched_yield (hundreds of oversubscribed threads: a futex round trip
// per collective would dominate the run time).  Acquire/release on `gen` orders everything before the barrier in any thread
// before everything after it in every thread, which is also what ThreadSanitizer sees.
struct Barrier {
  std::atomic<int> count{0};
  std::atomic<unsigned> gen{0};
  int n = 0;
  // returns false on time-out
  bool arrive_and_wait() {
    const unsigned g = gen.load(std::memory_order_acquire);
    if (count.fetch_add(1, std::memory_order_acq_rel) + 1 == n) {
      count.store(0, std::memory_order_relaxed);
      gen.store(g + 1, std::memory_order_release);
      return true;
    }
    const auto t0 = std::chrono::steady_clock::now();
    for (unsigned spins = 0; gen.load(std::memory_order_acquire) == g; spins++) {
      std::this_thread::yield();
      if ((spins & 0xfff) == 0xfff && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(wait_seconds())) return false;
    }
    return true;
  }
  void wait(const char* what) {
    if (!arrive_and_wait()) die(what);
  }
};

struct Warp {
  Barrier bar;
  uint64_t slot[2][32][2];
  // post-mortem of a collective that not every lane reached: where each lane last entered one, and whether it has left the kernel
  void* where[32][8] = {};
  int depth[32] = {};
  int state[32] = {};  // 0 running, 1 inside a collective, 2 exited
  void report() {
    for (int l = 0; l < 32; l++) {
      std::fprintf(stderr, "simt:   lane %2d %s", l, state[l] == 2 ? "exited " : (state[l] == 1 ? "waiting" : "running"));
      for (int k = 1; k < depth[l]; k++) {
        Dl_info info;
        if (dladdr(where[l][k], &info) && info.dli_fbase) std::fprintf(stderr, " +0x%zx", (size_t)((char*)where[l][k] - (char*)info.dli_fbase));
      }
      std::fprintf(stderr, "\n");
    }
  }
};

struct MBar {
  std::mutex m;
  std::condition_variable cv;
  int init_count = 0, pending = 0;
  long long tx = 0;
  unsigned phase = 0;
  void maybe_complete() {
    if (pending == 0 && tx == 0) {
      phase ^= 1u;
      pending = init_count;
      cv.notify_all();
    }
  }
};

// bar.sync / bar.arrive with a barrier id and a thread count (named barriers): arrivals are counted; the arrival that completes
// the count opens the barrier for the waiters of this generation
struct NamedBar {
  std::atomic<int> count{0};
  std::atomic<unsigned> gen{0};
  void arrive(int n) {
    if (count.fetch_add(1, std::memory_order_acq_rel) + 1 == n) {
      count.store(0, std::memory_order_relaxed);
      gen.fetch_add(1, std::memory_order_release);
    }
  }
  bool sync(int n) {
    const unsigned g = gen.load(std::memory_order_acquire);
    if (count.fetch_add(1, std::memory_order_acq_rel) + 1 == n) {
      count.store(0, std::memory_order_relaxed);
      gen.fetch_add(1, std::memory_order_release);
      return true;
    }
    const auto t0 = std::chrono::steady_clock::now();
    for (unsigned spins = 0; gen.load(std::memory_order_acquire) == g; spins++) {
      std::this_thread::yield();
      if ((spins & 0xfff) == 0xfff && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(wait_seconds())) return false;
    }
    return true;
  }
};

struct Block {
  Barrier bar;
  NamedBar named[16];
  std::atomic<int> votes[3] = {};  // __syncthreads_or: three rotating accumulators
  std::vector<std::unique_ptr<Warp>> warps;
  std::vector<unsigned char> dyn;
  std::mutex mbar_m;
  std::map<const void*, std::unique_ptr<MBar>> mbars;
  MBar& mbar(const void* p, bool create) {
    std::lock_guard<std::mutex> lk(mbar_m);
    auto it = mbars.find(p);
    if (it == mbars.end()) {
      if (!create) die("mbarrier used before mbarrier.init");
      it = mbars.emplace(p, std::make_unique<MBar>()).first;
    }
    return *it->second;
  }
};

struct ThreadCtx {
  Block* blk = nullptr;
  Warp* warp = nullptr;
  int lane = 0;
  unsigned op = 0, bop = 0;
};
inline thread_local ThreadCtx tc;

// warp barrier with the post-mortem bookkeeping (addresses are offsets into the test library: addr2line -e <lib> <offsets>)
inline void warp_wait(const char* what);

inline void* dyn_smem() { return tc.blk->dyn.data() + ((128 - (reinterpret_cast<uintptr_t>(tc.blk->dyn.data()) & 127)) & 127); }

}  // namespace simt

inline thread_local simt::Dim3 threadIdx, blockIdx;
inline simt::Dim3 blockDim, gridDim;

namespace simt {

// host threads that play the CUDA threads of a CTA; kept alive between launches (the lock-step engine launches thousands)
struct Pool {
  std::mutex m;
  std::condition_variable cv_start, cv_done;
  std::vector<std::thread> workers;
  unsigned epoch = 0, n_active = 0, remaining = 0;
  const std::function<void(unsigned)>* job = nullptr;
  bool stop = false;
  void worker(unsigned i) {
    unsigned seen = 0;
    std::unique_lock<std::mutex> lk(m);
    for (;;) {
      cv_start.wait(lk, [&] { return stop || epoch != seen; });
      if (stop) return;
      seen = epoch;
      if (i >= n_active) continue;
      const std::function<void(unsigned)>* j = job;
      lk.unlock();
      (*j)(i);
      lk.lock();
      if (--remaining == 0) cv_done.notify_one();
    }
  }
  void run(unsigned n, const std::function<void(unsigned)>& f) {
    std::unique_lock<std::mutex> lk(m);
    while (workers.size() < n) {
      const unsigned i = (unsigned)workers.size();
      workers.emplace_back([this, i] { worker(i); });
    }
    job = &f;
    n_active = remaining = n;
    epoch++;
    cv_start.notify_all();
    cv_done.wait(lk, [&] { return remaining == 0; });
  }
  ~Pool() {
    {
      std::lock_guard<std::mutex> lk(m);
      stop = true;
    }
    cv_start.notify_all();
    for (auto& t : workers) t.join();
  }
};
inline Pool& pool() {
  static Pool p;
  return p;
}

// run `body` as grid x block CUDA threads (1-D), one CTA at a time
inline void launch(dim3 grid, unsigned block, size_t dyn_bytes, const std::function<void()>& body) {
  if (block % 32) die("block size must be a multiple of 32");
  blockDim.x = block;
  gridDim.x = grid.x;
  gridDim.y = grid.y;
  for (unsigned b = 0; b < grid.x * grid.y; b++) {
    Block blk;
    blk.bar.n = (int)block;
    blk.dyn.assign(dyn_bytes + 256, 0xCD);  // poisoned: reading never-written shared memory shows up as garbage
    for (unsigned w = 0; w < block / 32; w++) {
      blk.warps.emplace_back(new Warp());
      blk.warps.back()->bar.n = 32;
    }
    pool().run(block, [&](unsigned t) {
      threadIdx.x = t;
      blockIdx.x = b % grid.x;
      blockIdx.y = b / grid.x;
      tc.blk = &blk;
      tc.warp = blk.warps[t / 32].get();
      tc.lane = (int)(t % 32);
      tc.op = 0;
      tc.bop = 0;
      body();
      tc.warp->state[tc.lane] = 2;
    });
  }
}

inline void warp_wait(const char* what) {
  Warp* w = tc.warp;
  const int l = tc.lane;
  w->depth[l] = backtrace(w->where[l], 8);
  w->state[l] = 1;
  if (!w->bar.arrive_and_wait()) {
    w->report();
    die(what);
  }
  w->state[l] = 0;
}

template <class T>
inline uint64_t bits(T v) {
  static_assert(sizeof(T) <= 8, "");
  uint64_t b = 0;
  std::memcpy(&b, &v, sizeof(T));
  return b;
}
template <class T>
inline T unbits(uint64_t b) {
  T v;
  std::memcpy(&v, &b, sizeof(T));
  return v;
}

// every lane posts (w0, w1); returns the slot array of the warp for this collective
inline const uint64_t (*exchange(uint64_t w0, uint64_t w1))[2] {
  Warp* w = tc.warp;
  const unsigned s = tc.op++ & 1u;
  w->slot[s][tc.lane][0] = w0;
  w->slot[s][tc.lane][1] = w1;
  warp_wait("warp collective: not all 32 lanes arrived (divergent collective, or lanes that left the kernel early)");
  return w->slot[s];
}

}  // namespace simt

inline void __syncthreads() { simt::tc.blk->bar.wait("__syncthreads: not all threads arrived"); }
namespace simt {
inline void named_bar_sync(int id, int nthreads) {
  if (!tc.blk->named[id].sync(nthreads)) die("bar.sync (named): not all threads arrived");
}
inline void named_bar_arrive(int id, int nthreads) { tc.blk->named[id].arrive(nthreads); }
}  // namespace simt
inline int __syncthreads_or(int pred) {
  simt::Block* b = simt::tc.blk;
  const unsigned n = simt::tc.bop++;
  // call n accumulates in slot n % 3 and clears slot (n + 1) % 3, which was last read in call n - 2: every thread finished that
  // call before it arrived at the barrier of call n - 1, which this thread has passed
  b->votes[(n + 1) % 3].store(0, std::memory_order_relaxed);
  if (pred) b->votes[n % 3].store(1, std::memory_order_relaxed);
  b->bar.wait("__syncthreads_or: not all threads arrived");
  return b->votes[n % 3].load(std::memory_order_relaxed);
}
inline void __syncwarp(unsigned = 0xffffffffu) { simt::warp_wait("__syncwarp: not all lanes arrived"); }
inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
#ifdef SIMT_MUTANT_EARLY_REFILL  // deliberately broken build (tests of the tests): the second arriver of eight is told it is the last
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST) + 6; }
#else
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
#endif

template <class T>
inline T __shfl_sync(unsigned, T v, int src) {
  auto s = simt::exchange(simt::bits(v), 0);
  return simt::unbits<T>(s[src & 31][0]);
}
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int mask) {
  auto s = simt::exchange(simt::bits(v), 0);
  return simt::unbits<T>(s[(simt::tc.lane ^ mask) & 31][0]);
}
template <class T>
inline T __ldcg(const T* p) { return *p; }
template <class T>
inline T __ldg(const T* p) { return *p; }

inline int __double2loint(double x) { return (int)(uint32_t)(simt::bits(x) & 0xffffffffu); }
inline int __double2hiint(double x) { return (int)(uint32_t)(simt::bits(x) >> 32); }
inline double __hiloint2double(int hi, int lo) { return simt::unbits<double>(((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo); }

namespace simt {

// mma.sync.aligned.m8n8k4.row.col.f64
inline void dmma(double& d0, double& d1, double a, double b) {
  auto s = exchange(bits(a), bits(b));
  const int lane = tc.lane, i = lane >> 2, j0 = 2 * (lane & 3);
  double r[2] = {d0, d1};
  for (int jj = 0; jj < 2; jj++) {
    const int j = j0 + jj;
    for (int k = 0; k < 4; k++) r[jj] = std::fma(unbits<double>(s[i * 4 + k][0]), unbits<double>(s[j * 4 + k][1]), r[jj]);
  }
  d0 = r[0];
  d1 = r[1];
}

inline void mbar_init(const void* bar, unsigned count) {
  MBar& m = tc.blk->mbar(bar, true);
  std::lock_guard<std::mutex> lk(m.m);
  m.init_count = m.pending = (int)count;
  m.tx = 0;
  m.phase = 0;
}
inline void mbar_expect_tx_arrive(const void* bar, unsigned bytes) {
  MBar& m = tc.blk->mbar(bar, false);
  std::lock_guard<std::mutex> lk(m.m);
  if (m.pending <= 0) die("mbarrier: more arrivals than the barrier was initialised for");
  m.tx += bytes;
  m.pending--;
  m.maybe_complete();
}
inline void mbar_complete_tx(const void* bar, unsigned bytes) {
  MBar& m = tc.blk->mbar(bar, false);
  std::lock_guard<std::mutex> lk(m.m);
  m.tx -= bytes;
  m.maybe_complete();
}
inline void mbar_wait(const void* bar, unsigned parity) {
#ifdef SIMT_MUTANT_SKIP_MBAR_WAIT  // deliberately broken build (tests of the tests): consumers do not wait for the bulk copies
  (void)bar;
  (void)parity;
  return;
#endif
  MBar& m = tc.blk->mbar(bar, false);
  std::unique_lock<std::mutex> lk(m.m);
  if (!m.cv.wait_for(lk, std::chrono::seconds(wait_seconds()), [&] { return (m.phase & 1u) != (parity & 1u); }))
    die("mbarrier.try_wait never succeeded (deadlock in the pipeline protocol)");
}
inline void tma_load_1d(void* smem_dst, const void* gsrc, unsigned bytes, const void* bar) {
  if ((bytes & 15u) || (reinterpret_cast<uintptr_t>(smem_dst) & 15u) || (reinterpret_cast<uintptr_t>(gsrc) & 15u))
    die("cp.async.bulk: size and both addresses must be multiples of 16 bytes");
  std::memcpy(smem_dst, gsrc, bytes);
  mbar_complete_tx(bar, bytes);
}

}  // namespace simt
