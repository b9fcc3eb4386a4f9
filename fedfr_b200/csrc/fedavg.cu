// FedAvg weighted parameter average (server.py:25-46) as one coalesced pointer-table reduction.
//
// out[e] = sum_i fl32(w_i) * fl32(src_i[e]), evaluated exactly like the reference's
// `tmp += weights[i] * models[i][name]`: an fp32 multiply rounded to nearest, then an fp32 add rounded
// to nearest, in client order (no FMA contraction) -- so results are bit-identical to the CPU path.
// int64 buffers (BatchNorm num_batches_tracked) take the same route through fp32.
//
// HBM bound: algorithmic bytes = (K + 1) * 4 per fp32 element.  Every thread keeps K independent
// 128-bit loads in flight (one per client) before the dependent add chain starts.
#include "common.cuh"

namespace pfc {

struct FedavgTable {          // device-side view of the segment table
  const void* const* src;     // [n_seg * K]
  float* const* out;          // [n_seg]
  const int64_t* len;         // [n_seg]
  const int32_t* dtype;       // [n_seg]
  const int64_t* blk_start;   // [n_seg + 1] first block of each segment
  const float* w;             // [K]
};

constexpr int kFedThreads = 256;
constexpr int kFedVecPerBlock = kFedThreads;          // one float4 per thread per block
constexpr int kFedMaxK = 64;

template <int kUnroll>
__device__ __forceinline__ void fedavg_f32_vec(const void* const* __restrict__ src, const float* __restrict__ w, int K, int64_t v,
                                               float4* __restrict__ out) {
  float4 acc;
  int i = 0;
  bool first = true;
  for (; i < K; i += kUnroll) {
    float4 x[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
      if (i + u < K) x[u] = ld_stream_f4(reinterpret_cast<const float4*>(src[i + u]) + v);
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      if (i + u < K) {
        const float wi = w[i + u];
        float4 t = make_float4(__fmul_rn(wi, x[u].x), __fmul_rn(wi, x[u].y), __fmul_rn(wi, x[u].z), __fmul_rn(wi, x[u].w));
        if (first) { acc = t; first = false; }
        else acc = make_float4(__fadd_rn(acc.x, t.x), __fadd_rn(acc.y, t.y), __fadd_rn(acc.z, t.z), __fadd_rn(acc.w, t.w));
      }
    }
  }
  st_stream_f4(out + v, acc);
}

__global__ void __launch_bounds__(kFedThreads) fedavg_kernel(FedavgTable tb, int n_seg, int K, int64_t total_blocks) {
  __shared__ float sw[kFedMaxK];
  __shared__ const void* sptr[kFedMaxK];
  if (threadIdx.x < K) sw[threadIdx.x] = tb.w[threadIdx.x];
  int cur_seg = -1;
  for (int64_t b = blockIdx.x; b < total_blocks; b += gridDim.x) {
    // locate the segment of block b (binary search over blk_start)
    int lo = 0, hi = n_seg - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (tb.blk_start[mid] <= b) lo = mid; else hi = mid - 1;
    }
    const int seg = lo;
    __syncthreads();
    if (seg != cur_seg) {
      if (threadIdx.x < K) sptr[threadIdx.x] = tb.src[(int64_t)seg * K + threadIdx.x];
      cur_seg = seg;
    }
    __syncthreads();
    const int64_t n = tb.len[seg];
    const int64_t local_blk = b - tb.blk_start[seg];
    float* out = tb.out[seg];
    if (tb.dtype[seg] == FEDAVG_F32) {
      const int64_t n_vec = n >> 2;
      const int64_t v = local_blk * kFedVecPerBlock + threadIdx.x;
      const bool aligned = (reinterpret_cast<uintptr_t>(out) & 15) == 0;   // sources checked on the host
      if (aligned && v < n_vec) fedavg_f32_vec<8>(sptr, sw, K, v, reinterpret_cast<float4*>(out));
      // scalar tail (and the unaligned fallback): elements [n_vec*4, n) of the segment, done by its last block
      const int64_t tail0 = aligned ? (n_vec << 2) : 0;
      const bool last_blk = (b + 1 == tb.blk_start[seg + 1]);
      if (!aligned || last_blk) {
        const int64_t e0 = aligned ? tail0 + threadIdx.x : local_blk * kFedVecPerBlock * 4 + threadIdx.x;
        const int64_t e1 = aligned ? n : min(n, (local_blk + 1) * (int64_t)kFedVecPerBlock * 4);
        for (int64_t e = e0; e < e1; e += kFedThreads) {
          float acc = __fmul_rn(sw[0], reinterpret_cast<const float*>(sptr[0])[e]);
          for (int i = 1; i < K; ++i) acc = __fadd_rn(acc, __fmul_rn(sw[i], reinterpret_cast<const float*>(sptr[i])[e]));
          out[e] = acc;
        }
      }
    } else {  // FEDAVG_I64: python_float * int64_tensor promotes to fp32 (server.py:32)
      const int64_t e0 = local_blk * kFedVecPerBlock * 4 + threadIdx.x;
      const int64_t e1 = min(n, (local_blk + 1) * (int64_t)kFedVecPerBlock * 4);
      for (int64_t e = e0; e < e1; e += kFedThreads) {
        float acc = __fmul_rn(sw[0], (float)reinterpret_cast<const long long*>(sptr[0])[e]);
        for (int i = 1; i < K; ++i) acc = __fadd_rn(acc, __fmul_rn(sw[i], (float)reinterpret_cast<const long long*>(sptr[i])[e]));
        out[e] = acc;
      }
    }
  }
}

__global__ void fedavg_blend_kernel(const float* __restrict__ old_fc, const float* __restrict__ aggr, float one_minus_p, float p, int64_t n,
                                    float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __fadd_rn(__fmul_rn(one_minus_p, old_fc[i]), __fmul_rn(p, aggr[i]));
}

static size_t align256(size_t x) { return (x + 255) / 256 * 256; }

}  // namespace pfc

using namespace pfc;

extern "C" {

size_t fedavg_table_bytes(int n_seg, int K) {
  if (n_seg < 0 || K < 0) return 0;
  return align256((size_t)n_seg * K * 8) + align256((size_t)n_seg * 8) * 2 + align256((size_t)n_seg * 4) + align256((size_t)(n_seg + 1) * 8) +
         align256((size_t)K * 4) + 256;
}

int fedavg_weighted_sum(const void* const* seg_src_host, void* const* seg_out_host, const int64_t* seg_len_host, const int32_t* seg_dtype_host,
                        int n_seg, const float* weights_host, int K, void* table_dev, size_t table_bytes, void* stream) {
  if (int rc = require_sm100()) return rc;
  PFC_REQUIRE(seg_src_host && seg_out_host && seg_len_host && seg_dtype_host && weights_host && table_dev, PFC_E_ARG, "fedavg_weighted_sum: null argument");
  PFC_REQUIRE(n_seg > 0 && K > 0 && K <= kFedMaxK, PFC_E_SHAPE, "fedavg_weighted_sum: need 1 <= K <= %d clients (got %d) and n_seg > 0", kFedMaxK, K);
  PFC_REQUIRE(table_bytes >= fedavg_table_bytes(n_seg, K), PFC_E_WORKSPACE, "fedavg_weighted_sum: table buffer too small");
  cudaStream_t st = as_stream(stream);
  // fp32 sources must be 16-byte aligned for the vector path (torch allocations are); verify on the host
  for (int s = 0; s < n_seg; ++s) {
    PFC_REQUIRE(seg_len_host[s] >= 0, PFC_E_ARG, "fedavg_weighted_sum: negative segment length");
    PFC_REQUIRE(seg_dtype_host[s] == FEDAVG_F32 || seg_dtype_host[s] == FEDAVG_I64, PFC_E_ARG, "fedavg_weighted_sum: unknown dtype code");
    if (seg_dtype_host[s] == FEDAVG_F32)
      for (int i = 0; i < K; ++i)
        PFC_REQUIRE((reinterpret_cast<uintptr_t>(seg_src_host[(size_t)s * K + i]) & 15) == 0 || seg_len_host[s] < 4, PFC_E_ARG,
                    "fedavg_weighted_sum: fp32 source %d of segment %d is not 16-byte aligned", i, s);
  }
  // block table (host, small): vector blocks cover 4*kFedVecPerBlock elements each
  static thread_local int64_t* blk_host = nullptr;
  static thread_local int blk_cap = 0;
  if (blk_cap < n_seg + 1) {
    if (blk_host) cudaFreeHost(blk_host);
    PFC_CUDA(cudaMallocHost(&blk_host, sizeof(int64_t) * (size_t)(n_seg + 1)));
    blk_cap = n_seg + 1;
  }
  int64_t total = 0;
  const int64_t per_blk = (int64_t)kFedVecPerBlock * 4;
  for (int s = 0; s < n_seg; ++s) {
    blk_host[s] = total;
    int64_t nb = (seg_len_host[s] + per_blk - 1) / per_blk;
    if (nb < 1) nb = 1;
    total += nb;
  }
  blk_host[n_seg] = total;
  char* p = reinterpret_cast<char*>(table_dev);
  FedavgTable tb;
  tb.src = reinterpret_cast<const void* const*>(p);
  PFC_CUDA(cudaMemcpyAsync(p, seg_src_host, (size_t)n_seg * K * 8, cudaMemcpyHostToDevice, st));
  p += align256((size_t)n_seg * K * 8);
  tb.out = reinterpret_cast<float* const*>(p);
  PFC_CUDA(cudaMemcpyAsync(p, seg_out_host, (size_t)n_seg * 8, cudaMemcpyHostToDevice, st));
  p += align256((size_t)n_seg * 8);
  tb.len = reinterpret_cast<const int64_t*>(p);
  PFC_CUDA(cudaMemcpyAsync(p, seg_len_host, (size_t)n_seg * 8, cudaMemcpyHostToDevice, st));
  p += align256((size_t)n_seg * 8);
  tb.dtype = reinterpret_cast<const int32_t*>(p);
  PFC_CUDA(cudaMemcpyAsync(p, seg_dtype_host, (size_t)n_seg * 4, cudaMemcpyHostToDevice, st));
  p += align256((size_t)n_seg * 4);
  tb.blk_start = reinterpret_cast<const int64_t*>(p);
  PFC_CUDA(cudaMemcpyAsync(p, blk_host, (size_t)(n_seg + 1) * 8, cudaMemcpyHostToDevice, st));
  p += align256((size_t)(n_seg + 1) * 8);
  tb.w = reinterpret_cast<const float*>(p);
  PFC_CUDA(cudaMemcpyAsync(p, weights_host, (size_t)K * 4, cudaMemcpyHostToDevice, st));
  int64_t grid = total;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (grid > cap) grid = cap;
  fedavg_kernel<<<(int)grid, kFedThreads, 0, st>>>(tb, n_seg, K, total);
  PFC_LAUNCH_CHECK();
  // blk_host is reused by the next call on this thread: make sure the copy above has been consumed
  PFC_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int fedavg_blend(const float* old_fc, const float* aggr, float one_minus_p, float p, int64_t n, float* out, void* stream) {
  if (int rc = require_sm100()) return rc;
  PFC_REQUIRE(old_fc && aggr && out && n >= 0, PFC_E_ARG, "fedavg_blend: bad argument");
  if (n == 0) return 0;
  int64_t blocks = (n + 255) / 256;
  if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
  fedavg_blend_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(old_fc, aggr, one_minus_p, p, n, out);
  PFC_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
