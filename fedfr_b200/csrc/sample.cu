// Sampled negative-centre selection (partial_fc.py:89-106) as an integer-exact index kernel chain.
//
//   index = sort(topk(perm with perm[positive] = 2.0, k = num_sample).indices)      (num_sample >= #positive)
//   index = positive                                                              (otherwise)
//   label[i] <- searchsorted(index, label[i])  for owned labels
//
// fp32 values in [0,1] U {2.0} are non-negative, so their bit patterns order like unsigned integers:
// the k-th largest is found with a 4 x 8-bit radix select (multi-block histograms, the last block to
// finish picks the bin), then one ordered stream compaction emits "all entries > kth, then entries
// == kth in ascending index order until k ids are out" -- the ids leave already sorted, so the
// reference's topk + sort collapses into select + compact.  Positives carry the value 2.0 (top byte
// 0x40, unreachable for torch.rand output), so the first histogram also counts them.
#include "common.cuh"

namespace pfc {

struct SampleState {
  unsigned int hist[256];
  unsigned int prefix;        // selected high bits of the k-th value so far
  unsigned int k_remaining;   // rank still to resolve inside the current prefix bucket
  unsigned int done;          // 1: kth is final (positives-only mode or k == 0)
  unsigned int kth_bits;      // final threshold
  unsigned int need_eq;       // how many entries == kth to take
  unsigned int take_none;     // k == 0
  unsigned int blocks_done;
  unsigned int n_pos;
  unsigned long long n_index;
};

constexpr int kSampleThreads = 256;
constexpr int kMaxCompactBlocks = 1024;

__global__ void sample_init_kernel(const int64_t* __restrict__ label, int64_t n_label, float* __restrict__ perm,
                                   int64_t num_local, SampleState* st) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (blockIdx.x == 0) {
    if (threadIdx.x < 256) st->hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) {
      st->prefix = 0; st->k_remaining = 0; st->done = 0; st->kth_bits = 0; st->need_eq = 0;
      st->take_none = 0; st->blocks_done = 0; st->n_pos = 0; st->n_index = 0;
    }
  }
  if (i < n_label) {
    const int64_t y = label[i];
    if (y >= 0 && y < num_local) perm[y] = 2.0f;      // perm[positive] = 2.0 (partial_fc.py:96)
  }
}

// One radix pass over byte `pass` (3 = most significant).  The last block selects the bin.
__global__ void __launch_bounds__(kSampleThreads) sample_radix_kernel(const float* __restrict__ perm, int64_t num_local,
                                                                      unsigned int num_sample, int pass, SampleState* st) {
  __shared__ unsigned int sh[256];
  __shared__ bool is_last;
  if (st->done) return;
  sh[threadIdx.x] = 0;
  __syncthreads();
  const unsigned int shift = pass * 8;
  const unsigned int hi_mask = pass == 3 ? 0u : (0xffffffffu << (shift + 8));
  const unsigned int prefix = st->prefix;
  const unsigned int* bits = reinterpret_cast<const unsigned int*>(perm);
  // warp-aggregated counting: torch.rand values crowd into a handful of top bytes (half of them share 0x3f), so one
  // shared-memory atomic per element serialises 32 ways; lanes with the same digit elect one to add their count
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n_iter = (num_local + stride - 1) / stride;          // the same for every thread: match is warp-convergent
  for (int64_t it = 0; it < n_iter; ++it) {
    const int64_t i = first + it * stride;
    unsigned int key = 0x100u + (unsigned int)lane;                   // not counted: a key of its own
    if (i < num_local) {
      const unsigned int v = bits[i];
      if ((v & hi_mask) == prefix) key = (v >> shift) & 0xffu;
    }
    const unsigned int peers = __match_any_sync(0xffffffffu, key);
    if (key < 0x100u && lane == __ffs((int)peers) - 1) atomicAdd(&sh[key], (unsigned int)__popc(peers));
  }
  __syncthreads();
  if (sh[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], sh[threadIdx.x]);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(&st->blocks_done, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (threadIdx.x == 0) {
    volatile unsigned int* h = st->hist;
    unsigned int k = st->k_remaining;
    if (pass == 3) {
      const unsigned int n_pos = h[0x40];
      st->n_pos = n_pos;
      if (num_sample < n_pos) {                 // index = positive (partial_fc.py:100-101)
        st->kth_bits = 0x40000000u; st->need_eq = n_pos; st->done = 1; st->n_index = n_pos;
      } else if (num_sample == 0) {
        st->take_none = 1; st->done = 1; st->kth_bits = 0xffffffffu; st->need_eq = 0; st->n_index = 0;
      } else {
        k = num_sample;
        st->n_index = num_sample;
      }
    }
    if (!st->done) {
      unsigned int cum = 0;
      int b = 255;
      for (; b > 0; --b) {
        if (cum + h[b] >= k) break;
        cum += h[b];
      }
      st->prefix = prefix | ((unsigned int)b << shift);
      st->k_remaining = k - cum;
      if (pass == 0) { st->kth_bits = st->prefix; st->need_eq = k - cum; }
    }
    st->blocks_done = 0;
  }
  __syncthreads();
  st->hist[threadIdx.x] = 0;
}

// Per-block counts of (> kth, == kth) over a contiguous chunk.
__global__ void __launch_bounds__(kSampleThreads) sample_count_kernel(const float* __restrict__ perm, int64_t num_local, int64_t chunk,
                                                                      const SampleState* st, unsigned int* __restrict__ blk_gt,
                                                                      unsigned int* __restrict__ blk_eq) {
  __shared__ unsigned int s_gt, s_eq;
  if (threadIdx.x == 0) { s_gt = 0; s_eq = 0; }
  __syncthreads();
  const unsigned int kth = st->kth_bits;
  const unsigned int* bits = reinterpret_cast<const unsigned int*>(perm);
  const int64_t lo = (int64_t)blockIdx.x * chunk;
  int64_t hi = lo + chunk;
  if (hi > num_local) hi = num_local;
  unsigned int gt = 0, eq = 0;
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const unsigned int v = bits[i];
    gt += v > kth;
    eq += v == kth;
  }
  for (int o = 16; o > 0; o >>= 1) { gt += __shfl_xor_sync(0xffffffffu, gt, o); eq += __shfl_xor_sync(0xffffffffu, eq, o); }
  if ((threadIdx.x & 31) == 0) { atomicAdd(&s_gt, gt); atomicAdd(&s_eq, eq); }
  __syncthreads();
  if (threadIdx.x == 0) { blk_gt[blockIdx.x] = s_gt; blk_eq[blockIdx.x] = s_eq; }
}

// Exclusive scan of the per-block counts (one block of kMaxCompactBlocks threads, one count pair per thread).
__global__ void __launch_bounds__(kMaxCompactBlocks) sample_scan_kernel(int n_blocks, unsigned int* blk_gt, unsigned int* blk_eq) {
  __shared__ unsigned int wsum_g[32], wsum_e[32];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const unsigned int g = t < n_blocks ? blk_gt[t] : 0u, e = t < n_blocks ? blk_eq[t] : 0u;
  unsigned int ig = g, ie = e;                                        // inclusive scan inside the warp
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int a = __shfl_sync(0xffffffffu, ig, lane >= o ? lane - o : lane);
    const unsigned int b = __shfl_sync(0xffffffffu, ie, lane >= o ? lane - o : lane);
    if (lane >= o) { ig += a; ie += b; }
  }
  if (lane == 31) { wsum_g[w] = ig; wsum_e[w] = ie; }
  __syncthreads();
  if (w == 0) {                                                       // scan of the warp totals
    unsigned int sg = wsum_g[lane], se = wsum_e[lane];
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int a = __shfl_sync(0xffffffffu, sg, lane >= o ? lane - o : lane);
      const unsigned int b = __shfl_sync(0xffffffffu, se, lane >= o ? lane - o : lane);
      if (lane >= o) { sg += a; se += b; }
    }
    wsum_g[lane] = sg; wsum_e[lane] = se;
  }
  __syncthreads();
  const unsigned int base_g = w > 0 ? wsum_g[w - 1] : 0u, base_e = w > 0 ? wsum_e[w - 1] : 0u;
  if (t < n_blocks) { blk_gt[t] = base_g + ig - g; blk_eq[t] = base_e + ie - e; }
}

// Ordered compaction: ids leave in ascending order.
__global__ void __launch_bounds__(kSampleThreads) sample_compact_kernel(const float* __restrict__ perm, int64_t num_local, int64_t chunk,
                                                                        const SampleState* st, const unsigned int* __restrict__ blk_gt,
                                                                        const unsigned int* __restrict__ blk_eq, int64_t* __restrict__ index_out) {
  __shared__ unsigned int warp_gt[kSampleThreads / 32], warp_eq[kSampleThreads / 32];
  __shared__ unsigned int run_gt, run_eq;
  const unsigned int kth = st->kth_bits;
  const unsigned int need_eq = st->need_eq;
  if (st->take_none) return;
  const unsigned int* bits = reinterpret_cast<const unsigned int*>(perm);
  const int64_t lo = (int64_t)blockIdx.x * chunk;
  int64_t hi = lo + chunk;
  if (hi > num_local) hi = num_local;
  if (threadIdx.x == 0) { run_gt = blk_gt[blockIdx.x]; run_eq = blk_eq[blockIdx.x]; }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int64_t base = lo; base < hi; base += blockDim.x) {
    const int64_t i = base + threadIdx.x;
    unsigned int v = 0;
    bool is_gt = false, is_eq = false;
    if (i < hi) { v = bits[i]; is_gt = v > kth; is_eq = v == kth; }
    const unsigned int m_gt = __ballot_sync(0xffffffffu, is_gt), m_eq = __ballot_sync(0xffffffffu, is_eq);
    if (lane == 0) { warp_gt[wid] = __popc(m_gt); warp_eq[wid] = __popc(m_eq); }
    __syncthreads();
    unsigned int off_gt = run_gt, off_eq = run_eq;
    for (int w = 0; w < wid; ++w) { off_gt += warp_gt[w]; off_eq += warp_eq[w]; }
    const unsigned int lt = (1u << lane) - 1u;
    const unsigned int my_gt = off_gt + __popc(m_gt & lt);      // # of > entries before me (global)
    const unsigned int my_eq = off_eq + __popc(m_eq & lt);      // # of == entries before me (global)
    const unsigned int eq_before = my_eq < need_eq ? my_eq : need_eq;
    if (is_gt) index_out[my_gt + eq_before] = i;
    else if (is_eq && my_eq < need_eq) index_out[my_gt + my_eq] = i;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int a = 0, b = 0;
      for (int w = 0; w < kSampleThreads / 32; ++w) { a += warp_gt[w]; b += warp_eq[w]; }
      run_gt += a; run_eq += b;
    }
    __syncthreads();
  }
}

// label[i] <- searchsorted(index, label[i])  (partial_fc.py:104); also publishes n_index.
__global__ void sample_relabel_kernel(int64_t* __restrict__ label, int64_t n_label, const int64_t* __restrict__ index,
                                      const SampleState* st, int64_t* n_index_out) {
  const int64_t n = (int64_t)st->n_index;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && n_index_out) *n_index_out = n;
  if (i >= n_label) return;
  const int64_t y = label[i];
  if (y < 0) return;
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (index[mid] < y) lo = mid + 1; else hi = mid;
  }
  label[i] = lo;
}

__global__ void remap_labels_kernel(const int64_t* __restrict__ in, int64_t n, int64_t class_start, int64_t num_local, int64_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int64_t y = in[i];
    out[i] = (y >= class_start && y < class_start + num_local) ? y - class_start : -1;
  }
}

static int compact_blocks(int64_t num_local, int64_t* chunk) {
  int64_t nb = (num_local + 2047) / 2048;
  if (nb > kMaxCompactBlocks) nb = kMaxCompactBlocks;
  if (nb < 1) nb = 1;
  int64_t c = (num_local + nb - 1) / nb;
  c = (c + kSampleThreads - 1) / kSampleThreads * kSampleThreads;
  *chunk = c;
  return (int)((num_local + c - 1) / c > 0 ? (num_local + c - 1) / c : 1);
}

}  // namespace pfc

using namespace pfc;

extern "C" {

int pfc_remap_labels(const int64_t* label_in, int64_t n, int64_t class_start, int64_t num_local, int64_t* label_out, void* stream) {
  if (int rc = require_sm100()) return rc;
  PFC_REQUIRE(label_in && label_out && n >= 0, PFC_E_ARG, "pfc_remap_labels: bad argument");
  if (n == 0) return 0;
  remap_labels_kernel<<<(int)((n + 255) / 256), 256, 0, as_stream(stream)>>>(label_in, n, class_start, num_local, label_out);
  PFC_LAUNCH_CHECK();
  return 0;
}

size_t pfc_sample_workspace_bytes(int64_t num_local) {
  (void)num_local;
  return sizeof(SampleState) + 2 * kMaxCompactBlocks * sizeof(unsigned int) + 256;
}

int pfc_sample_index(int64_t* label, int64_t n_label, float* perm, int64_t num_local, int64_t num_sample, int64_t* index_out,
                     int64_t* n_index_out, void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = require_sm100()) return rc;
  NvtxScope nvtx_scope("pfc_sample_index");
  PFC_REQUIRE(label && perm && index_out && workspace && n_label >= 0 && num_local > 0 && num_sample >= 0, PFC_E_ARG,
              "pfc_sample_index: bad argument");
  PFC_REQUIRE(num_local < (1ll << 31) && num_sample <= num_local, PFC_E_SHAPE, "pfc_sample_index: num_local/num_sample out of range");
  PFC_REQUIRE(workspace_bytes >= pfc_sample_workspace_bytes(num_local), PFC_E_WORKSPACE, "pfc_sample_index: workspace too small");
  cudaStream_t st = as_stream(stream);
  auto* state = reinterpret_cast<SampleState*>(workspace);
  auto* blk_gt = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(workspace) + ((sizeof(SampleState) + 255) / 256) * 256);
  auto* blk_eq = blk_gt + kMaxCompactBlocks;
  const int init_blocks = (int)((n_label > 256 ? n_label : 256) + 255) / 256;
  sample_init_kernel<<<init_blocks, 256, 0, st>>>(label, n_label, perm, num_local, state);
  PFC_LAUNCH_CHECK();
  int64_t hb = (num_local + kSampleThreads * 8 - 1) / (kSampleThreads * 8);
  if (hb > 2 * sm_count()) hb = 2 * sm_count();
  if (hb < 1) hb = 1;
  for (int pass = 3; pass >= 0; --pass) {
    sample_radix_kernel<<<(int)hb, kSampleThreads, 0, st>>>(perm, num_local, (unsigned int)num_sample, pass, state);
    PFC_LAUNCH_CHECK();
  }
  int64_t chunk = 0;
  const int nb = compact_blocks(num_local, &chunk);
  sample_count_kernel<<<nb, kSampleThreads, 0, st>>>(perm, num_local, chunk, state, blk_gt, blk_eq);
  PFC_LAUNCH_CHECK();
  sample_scan_kernel<<<1, kMaxCompactBlocks, 0, st>>>(nb, blk_gt, blk_eq);
  PFC_LAUNCH_CHECK();
  sample_compact_kernel<<<nb, kSampleThreads, 0, st>>>(perm, num_local, chunk, state, blk_gt, blk_eq, index_out);
  PFC_LAUNCH_CHECK();
  const int rb = (int)((n_label > 1 ? n_label : 1) + 255) / 256;
  sample_relabel_kernel<<<rb, 256, 0, st>>>(label, n_label, index_out, state, n_index_out);
  PFC_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
