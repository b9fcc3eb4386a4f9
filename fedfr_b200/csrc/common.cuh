// Shared helpers for the fedfr_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/fedfr_b200.h"

namespace pfc {

void set_error(const char* fmt, ...);
extern long long g_launch_count;      // kernels launched by this library (bench.py reports it)

// optional per-phase device timing (bench.py roofline): phases are bracketed by events on the launch stream
enum { PH_NORMALIZE = 0, PH_FWD = 1, PH_GRAD = 2, PH_DX = 3, PH_DW = 4, PH_COUNT = 5 };
void prof_begin(int phase, cudaStream_t st);
void prof_end(int phase, cudaStream_t st);
bool prof_enabled();

#define PFC_REQUIRE(cond, code, ...)                \
  do {                                              \
    if (!(cond)) {                                  \
      ::pfc::set_error(__VA_ARGS__);                \
      return (code);                                \
    }                                               \
  } while (0)

#define PFC_CUDA(call)                                                                 \
  do {                                                                                 \
    cudaError_t e__ = (call);                                                          \
    if (e__ != cudaSuccess) {                                                          \
      (void)cudaGetLastError();                                                        \
      ::pfc::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return (int)e__;                                                                 \
    }                                                                                  \
  } while (0)

#define PFC_LAUNCH_CHECK()                                                             \
  do {                                                                                 \
    ::pfc::g_launch_count++;                                                           \
    cudaError_t e__ = cudaGetLastError();                                              \
    if (e__ != cudaSuccess) {                                                          \
      ::pfc::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
      return (int)e__;                                                                 \
    }                                                                                  \
  } while (0)

// NVTX range around a C-ABI entry (host side, for nsys / ncu timelines).  Off unless FEDFR_NVTX=1 or pfc_set_nvtx(1).
void nvtx_push(const char* name);
void nvtx_pop();
struct NvtxScope {
  explicit NvtxScope(const char* name) { nvtx_push(name); }
  ~NvtxScope() { nvtx_pop(); }
};

// Checks once per process that the current device is compute capability 10.x.
int require_sm100();
int sm_count();

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// streaming 128-bit accesses that do not pollute L1
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_f4(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

// Margin applied to the target cosine (margin_kind: PFC_MARGIN_COSFACE / PFC_MARGIN_ARCFACE) and its slope
//   CosFace  losses.py:23-29   c - m                  slope 1
//   ArcFace  losses.py:38-45   cos(acos(c) + m)       slope sin(acos(c) + m) / sqrt(1 - c^2)   (autograd of acos_ / cos_)
// The cosine is clamped to [-1, 1] first (the reference would produce NaN outside; features and centres are unit vectors).
__device__ __forceinline__ float margin_cos(float c, float m, int kind) {
  if (kind == PFC_MARGIN_COSFACE) return c - m;
  c = fminf(fmaxf(c, -1.f), 1.f);
  return cosf(acosf(c) + m);
}
__device__ __forceinline__ float margin_slope(float c, float m, int kind) {
  if (kind == PFC_MARGIN_COSFACE) return 1.f;
  c = fminf(fmaxf(c, -1.f), 1.f);
  return sinf(acosf(c) + m) * rsqrtf(fmaxf(1.f - c * c, 1e-12f));
}

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

}  // namespace pfc
