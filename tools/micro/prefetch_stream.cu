// Micro-benchmark: how fast can a FEW warps per SM stream HBM (row sum of squares -> bf16 row out) with / without L2 bulk prefetch?
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ float4 ldf4(const float4* p) { float4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p)); return r; }
__device__ __forceinline__ void pf_bulk(const void* p, uint32_t bytes) { asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory"); }
__device__ __forceinline__ void pf_line(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p) : "memory"); }
template <int MODE, int ROWS>  // MODE 0 none, 1 bulk prefetch, 2 per-line prefetch
__global__ void __launch_bounds__(256) k(const float* __restrict__ w, int64_t n_rows, __nv_bfloat16* __restrict__ out, float* __restrict__ inv, int dist) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r0 = warp * ROWS; r0 < n_rows; r0 += n_warps * ROWS) {
    const int64_t rp = r0 + (int64_t)dist * n_warps * ROWS;
    if (MODE == 1 && lane < ROWS && rp + lane < n_rows) pf_bulk(w + (rp + lane) * 512, 2048);
    if (MODE == 2 && rp + ROWS <= n_rows) { for (int i = lane; i < ROWS * 16; i += 32) pf_line(w + rp * 512 + i * 32); }
    float4 v[ROWS][4];
#pragma unroll
    for (int k2 = 0; k2 < ROWS; ++k2)
#pragma unroll
      for (int i = 0; i < 4; ++i) v[k2][i] = ldf4(reinterpret_cast<const float4*>(w + (r0 + k2) * 512) + lane + i * 32);
#pragma unroll
    for (int k2 = 0; k2 < ROWS; ++k2) {
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) ss += v[k2][i].x * v[k2][i].x + v[k2][i].y * v[k2][i].y + v[k2][i].z * v[k2][i].z + v[k2][i].w * v[k2][i].w;
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const float rn = rsqrtf(fmaxf(ss, 1e-24f));
      if (lane == 0) inv[r0 + k2] = rn;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        __nv_bfloat162 a = __floats2bfloat162_rn(v[k2][i].x * rn, v[k2][i].y * rn), b = __floats2bfloat162_rn(v[k2][i].z * rn, v[k2][i].w * rn);
        uint2 pk = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
        *reinterpret_cast<uint2*>(out + (r0 + k2) * 512 + (lane + i * 32) * 4) = pk;
      }
    }
  }
}
template <int MODE, int ROWS> float run(const float* w, int64_t n, __nv_bfloat16* o, float* inv, int blocks, int dist) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE, ROWS><<<blocks, 256>>>(w, n, o, inv, dist);
  cudaEventRecord(a);
  for (int i = 0; i < 5; ++i) k<MODE, ROWS><<<blocks, 256>>>(w, n, o, inv, dist);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms / 5;
}
int main() {
  const int64_t n = 1000000 / 2 * 2; float* w; __nv_bfloat16* o; float* inv;
  cudaMalloc(&w, n * 2048); cudaMalloc(&o, n * 1024); cudaMalloc(&inv, n * 4); cudaMemset(w, 0x3c, n * 2048);
  const double gb = n * 3076.0 / 1e9;
  for (int bps : {1, 2, 4, 8}) {
    int blocks = 148 * bps;
    printf("blocks/SM %d: none %.0f GB/s", bps, gb / run<0, 2>(w, n, o, inv, blocks, 0) * 1e3);
    for (int d : {2, 4, 8, 16, 32}) printf(" | bulk d=%d %.0f", d, gb / run<1, 2>(w, n, o, inv, blocks, d) * 1e3);
    for (int d : {4, 16}) printf(" | line d=%d %.0f", d, gb / run<2, 2>(w, n, o, inv, blocks, d) * 1e3);
    printf("\n");
  }
  printf("err %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
