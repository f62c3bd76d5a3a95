// Test infrastructure: a CTA of the CUDA execution model on host threads, just enough of it to run the block-cooperative
// device functions of chromap_b200/csrc/ (shared arrays, __syncthreads, warp shuffles) unchanged on a machine without a GPU.
// One OS thread per CUDA thread; barriers are pthread barriers; a warp shuffle is an exchange through a per-warp buffer
// bracketed by two warp barriers.  Every lane of a warp must reach a shuffle (true of the functions under test: their scans
// are unconditional).  Not a performance tool.
#pragma once
#include <pthread.h>

#include <algorithm>
#include <cstdint>
#include <functional>
#include <thread>
#include <vector>

typedef unsigned long long u64;
typedef unsigned int u32;
typedef unsigned char u8;
#define __device__
#define __forceinline__ inline
#define __global__
using std::max;
using std::min;

struct EmuDim { int x; };
static thread_local EmuDim threadIdx, blockDim, blockIdx;

struct EmuCta {
  int nt = 0;
  pthread_barrier_t block;
  std::vector<pthread_barrier_t> warp;
  std::vector<u64> xch;  // [nt] exchange slots
};
static EmuCta *g_cta = nullptr;

static inline void __syncthreads() { pthread_barrier_wait(&g_cta->block); }
static inline void __syncwarp(unsigned = 0xffffffffu) { pthread_barrier_wait(&g_cta->warp[threadIdx.x >> 5]); }
template <typename T>
static inline T emu_shfl(T x, int src_lane) {  // value of lane src_lane of this warp (own value if out of range)
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  g_cta->xch[threadIdx.x] = (u64)x;
  pthread_barrier_wait(&g_cta->warp[w]);
  const int n_in_warp = std::min(32, g_cta->nt - w * 32);
  const T r = (src_lane >= 0 && src_lane < n_in_warp) ? (T)g_cta->xch[w * 32 + src_lane] : x;
  pthread_barrier_wait(&g_cta->warp[w]);
  (void)lane;
  return r;
}
template <typename T> static inline T __shfl_up_sync(unsigned, T x, int o) { const int l = threadIdx.x & 31; return emu_shfl(x, l - o >= 0 ? l - o : -1); }
template <typename T> static inline T __shfl_down_sync(unsigned, T x, int o) { const int l = threadIdx.x & 31; return emu_shfl(x, l + o < 32 ? l + o : -1); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T x, int o) { return emu_shfl(x, (threadIdx.x & 31) ^ o); }
template <typename T> static inline T __shfl_sync(unsigned, T x, int src) { return emu_shfl(x, src & 31); }

static inline unsigned __ballot_sync(unsigned, bool pred) {
  const int w = threadIdx.x >> 5;
  g_cta->xch[threadIdx.x] = pred ? 1u : 0u;
  pthread_barrier_wait(&g_cta->warp[w]);
  unsigned m = 0;
  const int n_in_warp = std::min(32, g_cta->nt - w * 32);
  for (int l = 0; l < n_in_warp; ++l) m |= (unsigned)(g_cta->xch[w * 32 + l] & 1u) << l;
  pthread_barrier_wait(&g_cta->warp[w]);
  return m;
}
template <typename T, typename Op>
static inline T emu_reduce(T x, Op op) {  // every lane gets op over the warp
  const int w = threadIdx.x >> 5;
  g_cta->xch[threadIdx.x] = (u64)(long long)x;
  pthread_barrier_wait(&g_cta->warp[w]);
  const int n_in_warp = std::min(32, g_cta->nt - w * 32);
  T r = (T)(long long)g_cta->xch[w * 32];
  for (int l = 1; l < n_in_warp; ++l) r = op(r, (T)(long long)g_cta->xch[w * 32 + l]);
  pthread_barrier_wait(&g_cta->warp[w]);
  return r;
}
static inline int __reduce_add_sync(unsigned, int x) { return emu_reduce<int>(x, [](int a, int b) { return a + b; }); }
static inline unsigned __reduce_add_sync(unsigned, unsigned x) { return emu_reduce<unsigned>(x, [](unsigned a, unsigned b) { return a + b; }); }
static inline int __reduce_max_sync(unsigned, int x) { return emu_reduce<int>(x, [](int a, int b) { return a > b ? a : b; }); }
static inline int __reduce_min_sync(unsigned, int x) { return emu_reduce<int>(x, [](int a, int b) { return a < b ? a : b; }); }
struct int4 { int x, y, z, w; };
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }

// run `body` as a CTA of nt threads (nt a multiple of 32)
static inline void emu_launch(int nt, const std::function<void()> &body) {
  EmuCta cta;
  cta.nt = nt;
  pthread_barrier_init(&cta.block, nullptr, (unsigned)nt);
  cta.warp.resize((size_t)(nt + 31) / 32);
  for (size_t w = 0; w < cta.warp.size(); ++w) pthread_barrier_init(&cta.warp[w], nullptr, (unsigned)std::min(32, nt - (int)w * 32));
  cta.xch.assign((size_t)nt, 0);
  g_cta = &cta;
  std::vector<std::thread> th;
  for (int t = 0; t < nt; ++t)
    th.emplace_back([&, t]() {
      threadIdx.x = t;
      blockDim.x = nt;
      body();
    });
  for (auto &x : th) x.join();
  pthread_barrier_destroy(&cta.block);
  for (auto &b : cta.warp) pthread_barrier_destroy(&b);
  g_cta = nullptr;
}
