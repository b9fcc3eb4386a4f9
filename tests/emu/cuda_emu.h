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

struct EmuDim3 {
  unsigned x = 1, y = 1, z = 1;
  EmuDim3() {}
  EmuDim3(unsigned x_, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
typedef EmuDim3 dim3;
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

template <class T> static inline T emu_shfl_exchange(T v, int src_lane_xor, int src_lane_abs) {
  static_assert(sizeof(T) <= 8, "shuffle of <= 64-bit values");
  static unsigned long long slots[64][32];
  const int t = (int)g_fibers[g_cur].tid.x, w = t / 32, l = t % 32;
  const int lanes = ((w + 1) * 32 <= (int)g_blockDim.x) ? 32 : (int)g_blockDim.x - w * 32;
  unsigned long long raw = 0;
  memcpy(&raw, &v, sizeof(T));
  slots[w][l] = raw;
  emu_warp_barrier(w, lanes);
  const int src = src_lane_abs >= 0 ? src_lane_abs : (l ^ src_lane_xor);
  raw = slots[w][src < lanes ? src : l];
  emu_warp_barrier(w, lanes);
  T r;
  memcpy(&r, &raw, sizeof(T));
  return r;
}
static inline unsigned __shfl_xor_sync(unsigned, unsigned v, int o) { return emu_shfl_exchange(v, o, -1); }
static inline int __shfl_xor_sync(unsigned, int v, int o) { return emu_shfl_exchange(v, o, -1); }
static inline double __shfl_xor_sync(unsigned, double v, int o) { return emu_shfl_exchange(v, o, -1); }
template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return emu_shfl_exchange(v, 0, src); }
static inline unsigned __ballot_sync(unsigned, bool pred) {
  unsigned m = 0;
  for (int b = 0; b < 32; ++b) m |= (emu_shfl_exchange<unsigned>(pred ? 1u : 0u, 0, b) & 1u) << b;   // lanes beyond the block read themselves
  const int t = (int)g_fibers[g_cur].tid.x, w = t / 32;
  const int lanes = ((w + 1) * 32 <= (int)g_blockDim.x) ? 32 : (int)g_blockDim.x - w * 32;
  return lanes == 32 ? m : (m & ((1u << lanes) - 1u));
}
static inline unsigned __match_any_sync(unsigned, unsigned key) {          // mask of the lanes of this warp holding the same key
  unsigned m = 0;
  const int t = (int)g_fibers[g_cur].tid.x, w = t / 32;
  const int lanes = ((w + 1) * 32 <= (int)g_blockDim.x) ? 32 : (int)g_blockDim.x - w * 32;
  for (int b = 0; b < 32; ++b) {
    const unsigned other = emu_shfl_exchange<unsigned>(key, 0, b);
    if (b < lanes && other == key) m |= 1u << b;
  }
  return m;
}
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline void __threadfence() {}
static inline void __syncwarp(unsigned = 0xffffffffu) {}

static void emu_trampoline() { g_body(); g_fibers[g_cur].done = true; }

template <class F> static void emu_launch(EmuDim3 grid, unsigned block, F body) {
  g_gridDim = grid; g_blockDim.x = block;
  g_body = body;
  // EMU_SCHEDULE=reverse runs the threads of a block (and the blocks of a grid) in descending order: a missing barrier
  // that ascending order happens to mask (writer has the lower index) is exposed by the other extreme
  static const bool reverse = getenv("EMU_SCHEDULE") && strcmp(getenv("EMU_SCHEDULE"), "reverse") == 0;
  for (unsigned bq = 0; bq < grid.x * grid.y; ++bq) {
    const unsigned b = reverse ? grid.x * grid.y - 1 - bq : bq;
    g_blockIdx.x = b % grid.x;
    g_blockIdx.y = b / grid.x;
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
      for (unsigned q = 0; q < block; ++q) {
        const unsigned t = reverse ? block - 1 - q : q;   // each fiber runs to its next rendezvous; the order decides who gets there first
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
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p += v; return o; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }

static inline int atomicAdd(int* p, int v) { int o = *p; *p += v; return o; }
static inline float atomicAdd(float* p, float v) { float o = *p; *p += v; return o; }

static inline int64_t min(int64_t a, int64_t b) { return a < b ? a : b; }
static inline int64_t max(int64_t a, int64_t b) { return a > b ? a : b; }

struct __nv_bfloat16 { uint16_t bits; };
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }

// vector types and streaming accessors
struct float4 { float x, y, z, w; };
struct uint2 { unsigned x, y; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }

// host runtime: "device" memory is host memory, streams are ignored
typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return 0; }
template <class T> static inline cudaError_t cudaMallocHost(T** p, size_t n) { *p = (T*)malloc(n); return *p ? 0 : 2; }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
typedef void* cudaEvent_t;
enum { cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (void*)1; return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }

#include "../../include/fedfr_b200.h"

namespace pfc {
static char g_emu_err[512];
static inline void set_error(const char* fmt, ...) { (void)fmt; snprintf(g_emu_err, sizeof(g_emu_err), "%s", fmt); }
static long long g_launch_count = 0;
static inline int require_sm100() { return 0; }
struct NvtxScope { explicit NvtxScope(const char*) {} };
enum { PH_NORMALIZE = 0, PH_FWD = 1, PH_GRAD = 2, PH_DX = 3, PH_DW = 4, PH_COUNT = 5 };
static inline void prof_begin(int, cudaStream_t) {}
static inline void prof_end(int, cudaStream_t) {}
static inline bool prof_enabled() { return false; }
static inline cudaStream_t as_stream(void* s) { return s; }
static inline float4 ld_stream_f4(const float4* p) { return *p; }
static inline void st_stream_f4(float4* p, const float4& v) { *p = v; }
#define PFC_REQUIRE(cond, code, ...) do { if (!(cond)) { ::pfc::set_error(__VA_ARGS__); return (code); } } while (0)
#define PFC_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return (int)e__; } while (0)
#define PFC_LAUNCH_CHECK() do { ::pfc::g_launch_count++; } while (0)
static inline int sm_count() { return 2; }
static inline float warp_max(float v) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
static inline unsigned emu_bf16_rn(float f) {            // round to nearest even, like __floats2bfloat162_rn
  unsigned u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fffu;
  return (u + 0x7fffu + ((u >> 16) & 1u)) >> 16;
}
static inline uint32_t pack_bf16x2(float lo, float hi) { return emu_bf16_rn(lo) | (emu_bf16_rn(hi) << 16); }
static inline float margin_cos(float c, float m, int kind) {          // same definitions as csrc/common.cuh
  if (kind == PFC_MARGIN_COSFACE) return c - m;
  c = fminf(fmaxf(c, -1.f), 1.f);
  return cosf(acosf(c) + m);
}
static inline float margin_slope(float c, float m, int kind) {
  if (kind == PFC_MARGIN_COSFACE) return 1.f;
  c = fminf(fmaxf(c, -1.f), 1.f);
  return sinf(acosf(c) + m) * rsqrtf(fmaxf(1.f - c * c, 1e-12f));
}
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
static inline float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
}  // namespace pfc
