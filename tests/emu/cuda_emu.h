// A thread-per-lane emulation of ONE CUDA thread block on the host, for `-m "not gpu"` tests of kernel logic
// (test infrastructure; nothing in the product includes it).  Every CUDA thread is an OS thread; __syncthreads,
// __syncwarp and the warp collectives (ballot, match, shuffles - full masks only) are pthread barriers plus an
// exchange buffer per warp.  Blocks of a launch run in blockIdx order, `window` of them at a time, which is how
// the hardware dispatches them; a block may therefore wait for an earlier one (decoupled look-back) but never
// for a later one.
#pragma once
#define __host__
#define __device__
#define __global__
#define __shared__
#include <cuda_runtime.h>      // vector types only (uint3, uint4, dim3); no CUDA call is made
#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif
#include <pthread.h>
#include <sched.h>

#include <atomic>
#include <cstdint>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

struct EmuCta {
  pthread_barrier_t cta_bar;
  pthread_barrier_t warp_bar[32];
  uint32_t xchg[32][32];
  std::vector<unsigned char> smem;
};
static thread_local EmuCta *emu_cta = nullptr;
static thread_local uint3 threadIdx, blockIdx;
static thread_local dim3 blockDim, gridDim;

inline void emu_yield() { sched_yield(); }
inline unsigned char *emu_dynamic_smem() { return emu_cta->smem.data(); }

inline void __syncthreads() { pthread_barrier_wait(&emu_cta->cta_bar); }
inline void __syncwarp(unsigned = 0xffffffffu) { pthread_barrier_wait(&emu_cta->warp_bar[threadIdx.x >> 5]); }
// every lane of the warp deposits a word, then reads all 32
inline void emu_exchange(uint32_t v, uint32_t out[32]) {
  EmuCta *c = emu_cta;
  const unsigned w = threadIdx.x >> 5, l = threadIdx.x & 31;
  c->xchg[w][l] = v;
  pthread_barrier_wait(&c->warp_bar[w]);
  std::memcpy(out, c->xchg[w], sizeof(uint32_t) * 32);
  pthread_barrier_wait(&c->warp_bar[w]);
}
inline uint32_t __ballot_sync(unsigned, bool p) {
  uint32_t o[32], m = 0;
  emu_exchange(p ? 1u : 0u, o);
  for (int i = 0; i < 32; i++) m |= (o[i] & 1u) << i;
  return m;
}
inline uint32_t __match_any_sync(unsigned, uint32_t v) {
  uint32_t o[32], m = 0;
  emu_exchange(v, o);
  for (int i = 0; i < 32; i++) m |= (uint32_t)(o[i] == v) << i;
  return m;
}
inline uint32_t __shfl_sync(unsigned, uint32_t v, unsigned src) { uint32_t o[32]; emu_exchange(v, o); return o[src & 31]; }
inline uint32_t __shfl_up_sync(unsigned, uint32_t v, unsigned d) {
  uint32_t o[32];
  emu_exchange(v, o);
  const unsigned l = threadIdx.x & 31;
  return l >= d ? o[l - d] : v;
}
inline uint32_t __shfl_xor_sync(unsigned, uint32_t v, unsigned x) { uint32_t o[32]; emu_exchange(v, o); return o[(threadIdx.x & 31) ^ x]; }
inline int __popc(uint32_t x) { return __builtin_popcount(x); }
inline int __ffs(uint32_t x) { return __builtin_ffs((int)x); }
inline int __clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }

// the helpers of b2_common.cuh that live under __CUDACC__
inline uint32_t lane_id() { return threadIdx.x & 31; }
inline uint32_t warp_id() { return threadIdx.x >> 5; }
inline uint32_t warp_incl_add(uint32_t v) {
  for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, v, o); if (lane_id() >= (uint32_t)o) v += t; }
  return v;
}

// Runs kernel() for every thread of every block of the grid; `window` blocks are resident at a time and blocks
// start in blockIdx order.
inline void emu_launch(unsigned grid, unsigned block, size_t smem_bytes, unsigned window, const std::function<void()> &kernel) {
  std::atomic<unsigned> next{0};
  auto worker = [&]() {
    for (;;) {
      const unsigned b = next.fetch_add(1);
      if (b >= grid) return;
      EmuCta cta;
      cta.smem.assign(smem_bytes + 16, 0xCD);              // poison: the kernel must initialise what it reads
      pthread_barrier_init(&cta.cta_bar, nullptr, block);
      for (unsigned w = 0; w < (block + 31) / 32; w++) pthread_barrier_init(&cta.warp_bar[w], nullptr, std::min(32u, block - 32 * w));
      std::vector<std::thread> th;
      th.reserve(block);
      for (unsigned t = 0; t < block; t++)
        th.emplace_back([&, t]() {
          emu_cta = &cta;
          threadIdx = uint3{t, 0, 0}; blockIdx = uint3{b, 0, 0};
          blockDim = dim3(block); gridDim = dim3(grid);
          kernel();
        });
      for (auto &x : th) x.join();
      pthread_barrier_destroy(&cta.cta_bar);
      for (unsigned w = 0; w < (block + 31) / 32; w++) pthread_barrier_destroy(&cta.warp_bar[w]);
    }
  };
  // blocks start in order because `next` is taken in order; a worker that holds block b runs it to completion
  std::vector<std::thread> ws;
  for (unsigned i = 0; i < window; i++) ws.emplace_back(worker);
  for (auto &x : ws) x.join();
}

inline bool __any_sync(unsigned, bool p) { return __ballot_sync(0xffffffffu, p) != 0u; }
inline uint32_t atomicOr(uint32_t *p, uint32_t v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
inline uint32_t atomicAdd(uint32_t *p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
using std::max;
using std::min;
