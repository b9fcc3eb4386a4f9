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
#include <string.h>

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
constexpr int kFedMaxK = 64;                          // clients per launch; more clients = more launches, same sum
constexpr int kFedUnaligned = 0x200;                 // internal flag in the dtype code: take the scalar path

// How the chain of a segment starts (the reference has both):
//   START_ZERO  FedPavg, server.py:30-32:       tmp = 0; tmp += w_0 x_0  ->  0.0f + t_0   (so -0.0 becomes +0.0)
//   START_TERM  FedAvg_on_FC, server.py:38:     aggr = x_0 * w_0          ->  t_0          (sign of zero kept)
//   START_OUT   clients 64.. of a K > 64 call:  the running sum of the previous launch, re-read from `out` (exact)
enum { START_ZERO = 0, START_TERM = 1, START_OUT = 2 };

template <int kUnroll>
__device__ __forceinline__ void fedavg_f32_vec(const void* const* __restrict__ src, const float* __restrict__ w, int K, int64_t v,
                                               float4* __restrict__ out, int start) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (start == START_OUT) acc = out[v];
  bool first = start == START_TERM;
  for (int i = 0; i < K; i += kUnroll) {
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

template <class T>
__device__ __forceinline__ float fedavg_scalar(const void* const* __restrict__ src, const float* __restrict__ w, int K, int64_t e, float prev, int start) {
  float acc = start == START_OUT ? prev : 0.f;
  bool first = start == START_TERM;
  for (int i = 0; i < K; ++i) {
    const float t = __fmul_rn(w[i], (float)reinterpret_cast<const T*>(src[i])[e]);
    if (first) { acc = t; first = false; }
    else acc = __fadd_rn(acc, t);
  }
  return acc;
}

// clients [k0, k0 + K) of a table whose rows hold k_total pointers
__global__ void __launch_bounds__(kFedThreads) fedavg_kernel(FedavgTable tb, int n_seg, int k_total, int k0, int K, int64_t total_blocks) {
  __shared__ float sw[kFedMaxK];
  __shared__ const void* sptr[kFedMaxK];
  if (threadIdx.x < K) sw[threadIdx.x] = tb.w[k0 + threadIdx.x];
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
      if (threadIdx.x < K) sptr[threadIdx.x] = tb.src[(int64_t)seg * k_total + k0 + threadIdx.x];
      cur_seg = seg;
    }
    __syncthreads();
    const int64_t n = tb.len[seg];
    const int64_t local_blk = b - tb.blk_start[seg];
    float* out = tb.out[seg];
    const int code = tb.dtype[seg];
    const int start = k0 > 0 ? START_OUT : ((code & FEDAVG_KEEP_FIRST_TERM) ? START_TERM : START_ZERO);
    if ((code & 0xff) == FEDAVG_F32) {
      const int64_t n_vec = n >> 2;
      const int64_t v = local_blk * kFedVecPerBlock + threadIdx.x;
      const bool aligned = !(code & kFedUnaligned);          // out and every source 16-byte aligned (checked on the host)
      if (aligned && v < n_vec) fedavg_f32_vec<8>(sptr, sw, K, v, reinterpret_cast<float4*>(out), start);
      // scalar tail (and the unaligned fallback): elements [n_vec*4, n) of the segment, done by its last block
      const int64_t tail0 = aligned ? (n_vec << 2) : 0;
      const bool last_blk = (b + 1 == tb.blk_start[seg + 1]);
      if (!aligned || last_blk) {
        const int64_t e0 = aligned ? tail0 + threadIdx.x : local_blk * kFedVecPerBlock * 4 + threadIdx.x;
        const int64_t e1 = aligned ? n : min(n, (local_blk + 1) * (int64_t)kFedVecPerBlock * 4);
        for (int64_t e = e0; e < e1; e += kFedThreads) out[e] = fedavg_scalar<float>(sptr, sw, K, e, start == START_OUT ? out[e] : 0.f, start);
      }
    } else {  // FEDAVG_I64: python_float * int64_tensor promotes to fp32 (server.py:32)
      const int64_t e0 = local_blk * kFedVecPerBlock * 4 + threadIdx.x;
      const int64_t e1 = min(n, (local_blk + 1) * (int64_t)kFedVecPerBlock * 4);
      for (int64_t e = e0; e < e1; e += kFedThreads) out[e] = fedavg_scalar<long long>(sptr, sw, K, e, start == START_OUT ? out[e] : 0.f, start);
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
  NvtxScope nvtx_scope("fedavg_weighted_sum");
  PFC_REQUIRE(seg_src_host && seg_out_host && seg_len_host && seg_dtype_host && weights_host && table_dev, PFC_E_ARG, "fedavg_weighted_sum: null argument");
  PFC_REQUIRE(n_seg > 0 && K > 0, PFC_E_SHAPE, "fedavg_weighted_sum: need K >= 1 clients (got %d) and n_seg > 0", K);
  PFC_REQUIRE(table_bytes >= fedavg_table_bytes(n_seg, K), PFC_E_WORKSPACE, "fedavg_weighted_sum: table buffer too small");
  cudaStream_t st = as_stream(stream);
  // host staging: a ring of pinned images of the whole device table (+ one event each).  The caller's arrays are copied
  // into the image before this call returns, so they may be reused at once, and consecutive calls do not wait for each
  // other's kernels (a slot is only waited for when the ring wraps around to it).
  struct Slot { char* host = nullptr; size_t cap = 0; cudaEvent_t ev = nullptr; };
  constexpr int kRing = 4;
  static thread_local Slot ring[kRing];
  static thread_local unsigned ring_pos = 0;
  Slot& slot = ring[ring_pos++ % kRing];
  const size_t need = fedavg_table_bytes(n_seg, K);
  if (slot.ev) PFC_CUDA(cudaEventSynchronize(slot.ev));
  else PFC_CUDA(cudaEventCreateWithFlags(&slot.ev, cudaEventDisableTiming));
  if (slot.cap < need) {
    if (slot.host) cudaFreeHost(slot.host);
    slot.host = nullptr; slot.cap = 0;
    PFC_CUDA(cudaMallocHost(&slot.host, need));
    slot.cap = need;
  }
  // image layout == device layout
  char* h = slot.host;
  const size_t off_out = align256((size_t)n_seg * K * 8), off_len = off_out + align256((size_t)n_seg * 8), off_code = off_len + align256((size_t)n_seg * 8);
  const size_t off_blk = off_code + align256((size_t)n_seg * 4), off_w = off_blk + align256((size_t)(n_seg + 1) * 8);
  int64_t* blk_host = reinterpret_cast<int64_t*>(h + off_blk);
  int32_t* code_host = reinterpret_cast<int32_t*>(h + off_code);
  memcpy(h, seg_src_host, (size_t)n_seg * K * 8);
  memcpy(h + off_out, seg_out_host, (size_t)n_seg * 8);
  memcpy(h + off_len, seg_len_host, (size_t)n_seg * 8);
  memcpy(h + off_w, weights_host, (size_t)K * 4);
  // the 128-bit path needs 16-byte aligned pointers (whole torch allocations are; views at odd offsets are not): a segment
  // with any unaligned source or output takes the scalar path instead
  for (int s = 0; s < n_seg; ++s) {
    PFC_REQUIRE(seg_len_host[s] >= 0, PFC_E_ARG, "fedavg_weighted_sum: negative segment length");
    const int base = seg_dtype_host[s] & 0xff;
    PFC_REQUIRE((base == FEDAVG_F32 || base == FEDAVG_I64) && (seg_dtype_host[s] & ~(0xff | FEDAVG_KEEP_FIRST_TERM)) == 0, PFC_E_ARG,
                "fedavg_weighted_sum: unknown dtype code");
    int code = seg_dtype_host[s];
    if (base == FEDAVG_F32) {
      bool ok = (reinterpret_cast<uintptr_t>(seg_out_host[s]) & 15) == 0;
      for (int i = 0; i < K && ok; ++i) ok = (reinterpret_cast<uintptr_t>(seg_src_host[(size_t)s * K + i]) & 15) == 0;
      PFC_REQUIRE((reinterpret_cast<uintptr_t>(seg_out_host[s]) & 3) == 0, PFC_E_ARG, "fedavg_weighted_sum: output %d is not 4-byte aligned", s);
      if (!ok) code |= kFedUnaligned;
    }
    code_host[s] = code;
  }
  // block table (host, small): vector blocks cover 4*kFedVecPerBlock elements each
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
  tb.out = reinterpret_cast<float* const*>(p + off_out);
  tb.len = reinterpret_cast<const int64_t*>(p + off_len);
  tb.dtype = reinterpret_cast<const int32_t*>(p + off_code);
  tb.blk_start = reinterpret_cast<const int64_t*>(p + off_blk);
  tb.w = reinterpret_cast<const float*>(p + off_w);
  PFC_CUDA(cudaMemcpyAsync(p, h, off_w + (size_t)K * 4, cudaMemcpyHostToDevice, st));
  int64_t grid = total;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (grid > cap) grid = cap;
  for (int k0 = 0; k0 < K; k0 += kFedMaxK) {               // K > 64: later launches continue the sums of the earlier ones (same order, same bits)
    const int kc = K - k0 < kFedMaxK ? K - k0 : kFedMaxK;
    fedavg_kernel<<<(int)grid, kFedThreads, 0, st>>>(tb, n_seg, K, k0, kc, total);
    PFC_LAUNCH_CHECK();
  }
  PFC_CUDA(cudaEventRecord(slot.ev, st));          // the image (and the device table) are free again once this has passed
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
