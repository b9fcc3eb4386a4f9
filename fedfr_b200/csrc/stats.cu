// Row-statistics plumbing of the distributed softmax (partial_fc.py:140-162).
//
// The forward kernels emit, per row, several partial (max, sum exp(z - max)) slots (one per CTA that
// touched the row).  merge_stats folds them into one triple per row; finalize_stats folds the triples
// of all ranks -- replacing all_reduce(MAX), all_reduce(SUM), all_reduce(SUM) at partial_fc.py:142,147,161
// by one all-gather -- and produces the loss:  -mean(log(max(exp(z_y - M) / S, 1e-30))).
#include "common.cuh"

namespace pfc {

// one warp per row: lanes stride over the partial slots
__global__ void __launch_bounds__(256) merge_stats_kernel(const float* __restrict__ part_max, const float* __restrict__ part_sum,
                                                          const float* __restrict__ target_logit, int n_part, int64_t n_rows,
                                                          float* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n_rows) return;
  float M = -INFINITY;
  for (int p = lane; p < n_part; p += 32)
    if (part_sum[(int64_t)p * n_rows + r] > 0.f) M = fmaxf(M, part_max[(int64_t)p * n_rows + r]);
  M = warp_max(M);
  float S = 0.f;
  for (int p = lane; p < n_part; p += 32) {
    const float l = part_sum[(int64_t)p * n_rows + r];
    if (l > 0.f) S += l * expf(part_max[(int64_t)p * n_rows + r] - M);
  }
  S = warp_sum(S);
  if (lane == 0) {
    stats[r * 3 + 0] = M;
    stats[r * 3 + 1] = S;
    stats[r * 3 + 2] = target_logit[r];
  }
}

// single block: rows strided over threads, block reduction for the loss
__global__ void __launch_bounds__(1024) finalize_stats_kernel(const float* __restrict__ g, int world, int64_t n_rows, float* __restrict__ row_max,
                                                              float* __restrict__ row_sum, float* __restrict__ loss_out) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int64_t r = threadIdx.x; r < n_rows; r += blockDim.x) {
    float M = -INFINITY;
    for (int w = 0; w < world; ++w) {
      const float* t = g + ((int64_t)w * n_rows + r) * 3;
      if (t[1] > 0.f) M = fmaxf(M, t[0]);
    }
    float S = 0.f, tz = 0.f;
    for (int w = 0; w < world; ++w) {
      const float* t = g + ((int64_t)w * n_rows + r) * 3;
      if (t[1] > 0.f) S += t[1] * expf(t[0] - M);
      tz += t[2];                                    // exactly one rank owns the target (others hold 0)
    }
    row_max[r] = M;
    row_sum[r] = S;
    const float p = expf(tz - M) / S;                // partial_fc.py:150,159
    acc += logf(fmaxf(p, 1e-30f));                   // partial_fc.py:162
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) *loss_out = -v / (float)n_rows;
  }
}

}  // namespace pfc

using namespace pfc;

extern "C" {

int pfc_merge_stats(const float* part_max, const float* part_sum, const float* target_logit, int n_partials, int64_t n_rows, float* stats,
                    void* stream) {
  if (int rc = require_sm100()) return rc;
  PFC_REQUIRE(part_max && part_sum && target_logit && stats && n_partials > 0 && n_rows >= 0, PFC_E_ARG, "pfc_merge_stats: bad argument");
  if (n_rows == 0) return 0;
  merge_stats_kernel<<<(int)((n_rows + 7) / 8), 256, 0, as_stream(stream)>>>(part_max, part_sum, target_logit, n_partials, n_rows, stats);
  PFC_LAUNCH_CHECK();
  return 0;
}

int pfc_finalize_stats(const float* gathered_stats, int world, int64_t n_rows, float* row_max, float* row_sum, float* loss_out, void* stream) {
  if (int rc = require_sm100()) return rc;
  PFC_REQUIRE(gathered_stats && row_max && row_sum && loss_out && world > 0 && n_rows > 0, PFC_E_ARG, "pfc_finalize_stats: bad argument");
  finalize_stats_kernel<<<1, 1024, 0, as_stream(stream)>>>(gathered_stats, world, n_rows, row_max, row_sum, loss_out);
  PFC_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
