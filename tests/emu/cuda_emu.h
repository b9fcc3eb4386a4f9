// Minimal single-OS-thread CUDA execution model for tests (tests/test_kernel_emulation.py): runs the SIMT kernels of
// fedfr_b200/csrc/{roc,bce_head}.cu on the CPU so their indexing, barriers and shuffles can be checked against the
// oracles without a GPU.  Every CUDA thread of a block is a ucontext fiber; __syncthreads() and __shfl_xor_sync() are
// rendezvous points (a fiber yields until its block / warp has arrived).  TEST INFRASTRUCTURE ONLY -- never shipped.
#pragma once
#include <ucontext.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

struct EmuDim3 { unsigned x = 1, y = 1, z = 1; };
struct EmuFiber { ucontext_t ctx; std::vector<char> stack; bool done = false; EmuDim3 tid; };

static EmuDim3 g_blockIdx, g_blockDim, g_gridDim;
static std::vector<EmuFiber> g_fibers;
static int g_cur = 0;
static ucontext_t g_main;
static std::function<void()> g_body;
static int g_bar_count = 0;
static unsigned g_bar_gen = 0;
static float g_shfl[64][32];
static int g_w_count[64];
static unsigned g_w_gen[64];

#define threadIdx (g_fibers[g_cur].tid)
#define blockIdx g_blockIdx
#define blockDim g_blockDim
#define gridDim g_gridDim
#define __global__
#define __device__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static

static inline void emu_yield() { swapcontext(&g_fibers[g_cur].ctx, &g_main); }

static inline void __syncthreads() {
  const unsigned gen = g_bar_gen;
  if (++g_bar_count == (int)g_blockDim.x) { g_bar_count = 0; ++g_bar_gen; }
  else while (g_bar_gen == gen) emu_yield();
}
static inline void emu_warp_barrier(int w, int lanes) {
  const unsigned gen = g_w_gen[w];
  if (++g_w_count[w] == lanes) { g_w_count[w] = 0; ++g_w_gen[w]; }
  else while (g_w_gen[w] == gen) emu_yield();
}
static inline float __shfl_xor_sync(unsigned, float v, int o) {
  const int t = (int)g_fibers[g_cur].tid.x, w = t / 32, l = t % 32;
  const int lanes = ((w + 1) * 32 <= (int)g_blockDim.x) ? 32 : (int)g_blockDim.x - w * 32;
  g_shfl[w][l] = v;
  emu_warp_barrier(w, lanes);
  const float r = g_shfl[w][l ^ o];
  emu_warp_barrier(w, lanes);
  return r;
}

static void emu_trampoline() { g_body(); g_fibers[g_cur].done = true; }

template <class F> static void emu_launch(unsigned grid, unsigned block, F body) {
  g_gridDim.x = grid; g_blockDim.x = block;
  g_body = body;
  for (unsigned b = 0; b < grid; ++b) {
    g_blockIdx.x = b;
    g_bar_count = 0;
    memset(g_w_count, 0, sizeof(g_w_count));
    g_fibers.assign(block, EmuFiber());
    for (unsigned t = 0; t < block; ++t) {
      EmuFiber& f = g_fibers[t];
      f.stack.resize(256 * 1024);
      f.tid.x = t;
      getcontext(&f.ctx);
      f.ctx.uc_stack.ss_sp = f.stack.data();
      f.ctx.uc_stack.ss_size = f.stack.size();
      f.ctx.uc_link = &g_main;
      makecontext(&f.ctx, emu_trampoline, 0);
    }
    for (long rounds = 0;; ++rounds) {
      bool any = false;
      for (unsigned t = 0; t < block; ++t) {
        if (g_fibers[t].done) continue;
        any = true;
        g_cur = (int)t;
        swapcontext(&g_main, &g_fibers[t].ctx);
      }
      if (!any) break;
      if (rounds > 100000000L) { fprintf(stderr, "cuda_emu: deadlock (divergent barrier?)\n"); abort(); }
    }
  }
}

// arithmetic with CUDA's names (compile with -ffp-contract=off so * and + stay separate roundings)
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p += v; return o; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }

namespace pfc {
static inline int sm_count() { return 2; }
static inline float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
}  // namespace pfc
