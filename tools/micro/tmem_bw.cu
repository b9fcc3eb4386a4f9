// Micro-benchmark (developer tool, not part of the library): tcgen05.ld throughput per SM as a function of the number of
// reading warps, and shared-memory store + TMA-store drain rate of an epilogue-like loop.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tmem_bw tools/micro/tmem_bw.cu && /tmp/tmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// every warp reads `cols` columns of its lane quadrant per round (x32 loads, `depth` loads in flight before a wait)
template <int DEPTH>
__global__ void tmem_read_kernel(int rounds, int cols_per_warp, long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * cols_per_warp) % 512;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
    for (int c = 0; c < cols_per_warp; c += 32 * DEPTH) {
      uint32_t v[DEPTH][32];
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) tmem_ld_x32(base + c + d * 32, v[d]);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int d = 0; d < DEPTH; ++d)
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= v[d][j];
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

// shared-memory store rate of an epilogue-like loop: every thread writes 16-byte chunks of its own 128-byte row
__global__ void smem_store_kernel(int rounds, long long* out, uint32_t* sink) {
  extern __shared__ uint8_t sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* box = sm + warp * 4096;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
#pragma unroll
    for (int q = 0; q < 8; ++q)
      *reinterpret_cast<uint4*>(box + lane * 128 + ((q ^ (lane & 7)) << 4)) = make_uint4(r, q, lane, warp);
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (sm[threadIdx.x] == 0xff && rounds < 0) sink[0] = 1;
}

int main() {
  long long* d_out;
  uint32_t* d_sink;
  cudaMalloc(&d_out, 1024 * sizeof(long long));
  cudaMalloc(&d_sink, 16);
  long long h[1024];
  const int rounds = 200;
  for (int warps : {4, 8, 16}) {
    const int cols = 512 / (warps / 4);
    for (int depth : {1, 2, 4}) {
      if (cols < 32 * depth || (depth == 4 && warps > 8)) continue;
      if (depth == 1) tmem_read_kernel<1><<<148, warps * 32>>>(rounds, cols, d_out, d_sink);
      if (depth == 2) tmem_read_kernel<2><<<148, warps * 32>>>(rounds, cols, d_out, d_sink);
      if (depth == 4) tmem_read_kernel<4><<<148, warps * 32>>>(rounds, cols, d_out, d_sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d_out, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
      const double bytes = (double)rounds * 128 * 512 * 4;      // the whole 256 KB of TMEM per round
      printf("{\"probe\": \"tmem_read\", \"warps\": %d, \"depth\": %d, \"cycles\": %lld, \"bytes_per_clk_per_sm\": %.1f}\n", warps, depth, h[0],
             bytes / (double)h[0]);
    }
  }
  for (int warps : {4, 8, 16}) {
    cudaFuncSetAttribute(smem_store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 4096);
    smem_store_kernel<<<148, warps * 32, 16 * 4096>>>(2000, d_out, d_sink);
    cudaDeviceSynchronize();
    cudaMemcpy(h, d_out, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
    printf("{\"probe\": \"smem_store_v4\", \"warps\": %d, \"cycles\": %lld, \"bytes_per_clk_per_sm\": %.1f}\n", warps, h[0],
           2000.0 * warps * 4096 / (double)h[0]);
  }
  return 0;
}
