// HBM-bound row kernels: normalise (+gather) -> bf16, cast, gather, scatter.
// One warp per row, 128-bit accesses; rows are 2 KB (E=512 fp32) so a warp owns 4 float4 per lane.
#include "common.cuh"
#include "rows_device.cuh"

namespace pfc {

// normalize(sub_weight) (partial_fc.py:127) fused with the sampled gather (partial_fc.py:105).
// Algorithmic bytes per row: read 4E, write 2E (bf16) + 4.
template <int kVecPerLane, int kRows>
__global__ void __launch_bounds__(256) normalize_rows_kernel(const float* __restrict__ w, const int64_t* __restrict__ index,
                                                             int64_t n_rows, int emb, __nv_bfloat16* __restrict__ out_bf16,
                                                             float* __restrict__ out_f32, float* __restrict__ inv_norm) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  normalize_rows_warp<kVecPerLane, kRows>(w, index, n_rows, emb, out_bf16, out_f32, inv_norm, warp, n_warps, lane);
}

__global__ void __launch_bounds__(256) cast_rows_kernel(const float4* __restrict__ x, int64_t n_vec, uint2* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = x[i];
    out[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
}

// sub_weight = weight[index]; sub_weight_mom = weight_mom[index]   (gather=true)
// weight[index] = sub_weight; weight_mom[index] = sub_weight_mom   (gather=false)
template <bool kGather>
__global__ void __launch_bounds__(256) move_rows2_kernel(float* __restrict__ big_a, float* __restrict__ big_b,
                                                         const int64_t* __restrict__ index, int64_t n_index, int emb,
                                                         float* __restrict__ small_a, float* __restrict__ small_b) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int nvec = emb >> 2;
  for (int64_t r = warp; r < n_index; r += n_warps) {
    const int64_t row = index[r];
    float4* ba = reinterpret_cast<float4*>(big_a + row * emb);
    float4* bb = big_b ? reinterpret_cast<float4*>(big_b + row * emb) : nullptr;
    float4* sa = reinterpret_cast<float4*>(small_a + r * emb);
    float4* sb = small_b ? reinterpret_cast<float4*>(small_b + r * emb) : nullptr;
    for (int c = lane; c < nvec; c += 32) {
      if (kGather) {
        st_stream_f4(sa + c, ld_stream_f4(ba + c));
        if (bb && sb) st_stream_f4(sb + c, ld_stream_f4(bb + c));
      } else {
        st_stream_f4(ba + c, ld_stream_f4(sa + c));
        if (bb && sb) st_stream_f4(bb + c, ld_stream_f4(sb + c));
      }
    }
  }
}

// losses.CosFace.forward on materialised logits (losses.py:23-29).
__global__ void cosface_margin_kernel(float* cosine, const int64_t* label, int64_t n_rows, int64_t n_classes, float m) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_rows) {
    int64_t y = label[i];
    if (y >= 0 && y < n_classes) cosine[i * n_classes + y] -= m;
  }
}
// losses.ArcFace.forward on materialised logits (losses.py:38-45), in place like the reference: acos_ over the whole matrix,
// + m on the target column of rows with label != -1, cos_, mul_(s).  One pass; |cosine| > 1 gives NaN as in the reference.
__global__ void arcface_dense_kernel(float* __restrict__ cosine, const int64_t* __restrict__ label, int64_t n_rows, int64_t n_classes, float s, float m) {
  const int64_t n = n_rows * n_classes;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / n_classes, c = i - r * n_classes;
    float th = acosf(cosine[i]);
    if (label[r] == c) th += m;
    cosine[i] = cosf(th) * s;
  }
}
__global__ void scale_kernel(const float* __restrict__ in, float s, int64_t n, float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = in[i] * s;
}

// Fused optimizer.step() + update() for the head: torch.optim.SGD (momentum, weight decay, dampening,
// nesterov) on the rows of the shard named by `index` (all rows when NULL), written in place -- for a sampled step this is
// `weight[index]`, `weight_mom[index]`, i.e. PartialFC.update() (partial_fc.py:113-116) never has to run.  The arithmetic
// follows torch's foreach kernels operation by operation (a + alpha * b is one fma):
//     g   = fma(wd, w, grad)            _foreach_add(grads, params, alpha = weight_decay)
//     buf = buf * momentum              _foreach_mul_(bufs, momentum)
//     buf = fma(1 - dampening, g, buf)  _foreach_add_(bufs, grads, alpha = 1 - dampening)
//     g   = fma(momentum, buf, g)       (nesterov only)
//     w   = fma(-lr, buf or g, w)       _foreach_add_(params, grads, alpha = -lr)
// Optionally emits the bf16 normalised rows + 1/norm of the UPDATED weights, which is what the next forward needs
// (saves that step's separate normalize pass over the shard).  One warp per row; HBM bound: 5 x 4E bytes per row.
template <int kVecPerLane>
__global__ void __launch_bounds__(256) sgd_rows_kernel(float* __restrict__ weight, float* __restrict__ mom, const float* __restrict__ grad,
                                                       const int64_t* __restrict__ index, int64_t n_rows, int emb, float lr, float momentum,
                                                       float one_minus_damp, float wd, int nesterov, __nv_bfloat16* __restrict__ w_hat,
                                                       float* __restrict__ inv_norm) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int nvec = emb >> 2;
  for (int64_t r = warp; r < n_rows; r += n_warps) {
    const int64_t dst = index ? index[r] : r;
    float4* pw = reinterpret_cast<float4*>(weight + dst * emb);
    float4* pm = reinterpret_cast<float4*>(mom + dst * emb);
    const float4* pg = reinterpret_cast<const float4*>(grad + r * emb);
    float4 w[kVecPerLane], b[kVecPerLane], g[kVecPerLane];
#pragma unroll
    for (int i = 0; i < kVecPerLane; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) { w[i] = ld_stream_f4(pw + c); b[i] = ld_stream_f4(pm + c); g[i] = ld_stream_f4(pg + c); }
    }
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < kVecPerLane; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
        float* wf = reinterpret_cast<float*>(&w[i]);
        float* bf = reinterpret_cast<float*>(&b[i]);
        float* gf = reinterpret_cast<float*>(&g[i]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float gg = wd != 0.f ? __fmaf_rn(wd, wf[k], gf[k]) : gf[k];
          float bb = __fmul_rn(bf[k], momentum);
          bb = __fmaf_rn(one_minus_damp, gg, bb);
          const float step = nesterov ? __fmaf_rn(momentum, bb, gg) : bb;
          wf[k] = __fmaf_rn(-lr, step, wf[k]);
          bf[k] = bb;
        }
        ss += w[i].x * w[i].x + w[i].y * w[i].y + w[i].z * w[i].z + w[i].w * w[i].w;      // same expression as normalize_rows_warp: same bits
        st_stream_f4(pw + c, w[i]);
        st_stream_f4(pm + c, b[i]);
      }
    }
    if (w_hat != nullptr || inv_norm != nullptr) {
      ss = warp_sum(ss);
      const float nrm = fmaxf(sqrtf(ss), 1e-12f);
      const float inv = 1.0f / nrm;
      if (lane == 0 && inv_norm) inv_norm[r] = inv;
      if (w_hat) {
#pragma unroll
        for (int i = 0; i < kVecPerLane; ++i) {
          const int c = lane + i * 32;
          if (c < nvec)
            *reinterpret_cast<uint2*>(w_hat + r * emb + c * 4) = make_uint2(pack_bf16x2(w[i].x * inv, w[i].y * inv), pack_bf16x2(w[i].z * inv, w[i].w * inv));
        }
      }
    }
  }
}

static int row_grid(int64_t n_rows) {
  int64_t blocks = (n_rows + 7) / 8;            // 8 warps per block
  int64_t cap = (int64_t)sm_count() * 8;        // 8 resident blocks of 256 threads per SM
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// blocks_per_sm > 0 caps the grid at that many blocks per SM and keeps two rows per warp in flight: the shape used when
// the kernel shares the SMs with a tensor-core kernel (pfc_normalize_fwd_stats).
int launch_normalize_rows(const float* w, const int64_t* index, int64_t n_rows, int emb, __nv_bfloat16* ob, float* of, float* inv_norm,
                          int blocks_per_sm, cudaStream_t st) {
  PFC_REQUIRE(w && n_rows >= 0 && emb > 0, PFC_E_ARG, "pfc_normalize_rows: bad argument");
  PFC_REQUIRE(emb % 4 == 0 && emb <= 2048, PFC_E_SHAPE, "pfc_normalize_rows: emb=%d must be a multiple of 4 and <= 2048", emb);
  if (n_rows == 0) return 0;
  const int vec_per_lane = (emb / 4 + 31) / 32;
  int grid = row_grid(n_rows);
  if (blocks_per_sm > 0) {
    int64_t blocks = (n_rows + 15) / 16;
    const int64_t cap = (int64_t)sm_count() * blocks_per_sm;
    grid = (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
    if (vec_per_lane <= 1) normalize_rows_kernel<1, 2><<<grid, 256, 0, st>>>(w, index, n_rows, emb, ob, of, inv_norm);
    else if (vec_per_lane <= 2) normalize_rows_kernel<2, 2><<<grid, 256, 0, st>>>(w, index, n_rows, emb, ob, of, inv_norm);
    else if (vec_per_lane <= 4) normalize_rows_kernel<4, 2><<<grid, 256, 0, st>>>(w, index, n_rows, emb, ob, of, inv_norm);
    else if (vec_per_lane <= 8) normalize_rows_kernel<8, 1><<<grid, 256, 0, st>>>(w, index, n_rows, emb, ob, of, inv_norm);
    else normalize_rows_kernel<16, 1><<<grid, 256, 0, st>>>(w, index, n_rows, emb, ob, of, inv_norm);
  } else {
    if (vec_per_lane <= 1) normalize_rows_kernel<1, 1><<<grid, 256, 0, st>>>(w, index, n_rows, emb, ob, of, inv_norm);
    else if (vec_per_lane <= 2) normalize_rows_kernel<2, 1><<<grid, 256, 0, st>>>(w, index, n_rows, emb, ob, of, inv_norm);
    else if (vec_per_lane <= 4) normalize_rows_kernel<4, 1><<<grid, 256, 0, st>>>(w, index, n_rows, emb, ob, of, inv_norm);
    else if (vec_per_lane <= 8) normalize_rows_kernel<8, 1><<<grid, 256, 0, st>>>(w, index, n_rows, emb, ob, of, inv_norm);
    else normalize_rows_kernel<16, 1><<<grid, 256, 0, st>>>(w, index, n_rows, emb, ob, of, inv_norm);
  }
  PFC_LAUNCH_CHECK();
  return 0;
}

}  // namespace pfc

using namespace pfc;

extern "C" {

int pfc_normalize_rows(const float* w, const int64_t* index, int64_t n_rows, int emb, void* w_hat_bf16, float* w_hat_f32,
                       float* inv_norm, void* stream) {
  if (int rc = require_sm100()) return rc;
  NvtxScope nvtx_scope("pfc_normalize_rows");
  cudaStream_t st = as_stream(stream);
  prof_begin(PH_NORMALIZE, st);
  if (int rc = launch_normalize_rows(w, index, n_rows, emb, reinterpret_cast<__nv_bfloat16*>(w_hat_bf16), w_hat_f32, inv_norm, 0, st)) return rc;
  prof_end(PH_NORMALIZE, st);
  return 0;
}

int pfc_sgd_step(float* weight, float* weight_mom, const float* grad, const int64_t* index, int64_t n_rows, int emb, float lr, float momentum,
                 float dampening, float weight_decay, int nesterov, void* w_hat_bf16, float* inv_norm, void* stream) {
  if (int rc = require_sm100()) return rc;
  NvtxScope nvtx_scope("pfc_sgd_step");
  PFC_REQUIRE(weight && weight_mom && grad && n_rows >= 0 && emb > 0, PFC_E_ARG, "pfc_sgd_step: bad argument");
  PFC_REQUIRE(emb % 4 == 0 && emb <= 2048, PFC_E_SHAPE, "pfc_sgd_step: emb=%d must be a multiple of 4 and <= 2048", emb);
  PFC_REQUIRE(!(index && (w_hat_bf16 || inv_norm)), PFC_E_ARG, "pfc_sgd_step: the normalised output is for unsampled steps (index == NULL)");
  if (n_rows == 0) return 0;
  const int vec_per_lane = (emb / 4 + 31) / 32;
  const int grid = row_grid(n_rows);
  cudaStream_t st = as_stream(stream);
  auto* wh = reinterpret_cast<__nv_bfloat16*>(w_hat_bf16);
  const float omd = 1.0f - dampening;
#define PFC_SGD(V) sgd_rows_kernel<V><<<grid, 256, 0, st>>>(weight, weight_mom, grad, index, n_rows, emb, lr, momentum, omd, weight_decay, nesterov, wh, inv_norm)
  if (vec_per_lane <= 1) PFC_SGD(1);
  else if (vec_per_lane <= 2) PFC_SGD(2);
  else if (vec_per_lane <= 4) PFC_SGD(4);
  else if (vec_per_lane <= 8) PFC_SGD(8);
  else PFC_SGD(16);
#undef PFC_SGD
  PFC_LAUNCH_CHECK();
  return 0;
}

int pfc_cast_rows_bf16(const float* x, int64_t n_rows, int emb, void* x_bf16, void* stream) {
  if (int rc = require_sm100()) return rc;
  PFC_REQUIRE(x && x_bf16 && n_rows >= 0 && emb > 0 && emb % 4 == 0, PFC_E_ARG, "pfc_cast_rows_bf16: bad argument");
  if (n_rows == 0) return 0;
  const int64_t n_vec = n_rows * emb / 4;
  int64_t blocks = (n_vec + 255) / 256;
  if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
  cast_rows_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(x), n_vec, reinterpret_cast<uint2*>(x_bf16));
  PFC_LAUNCH_CHECK();
  return 0;
}

int pfc_gather_rows2(const float* weight, const float* weight_mom, const int64_t* index, int64_t n_index, int emb,
                     float* sub_weight, float* sub_weight_mom, void* stream) {
  if (int rc = require_sm100()) return rc;
  NvtxScope nvtx_scope("pfc_gather_rows2");
  PFC_REQUIRE(weight && index && sub_weight && n_index >= 0 && emb > 0 && emb % 4 == 0, PFC_E_ARG, "pfc_gather_rows2: bad argument");
  if (n_index == 0) return 0;
  move_rows2_kernel<true><<<row_grid(n_index), 256, 0, as_stream(stream)>>>(const_cast<float*>(weight), const_cast<float*>(weight_mom), index,
                                                                             n_index, emb, sub_weight, sub_weight_mom);
  PFC_LAUNCH_CHECK();
  return 0;
}

int pfc_scatter_rows2(float* weight, float* weight_mom, const int64_t* index, int64_t n_index, int emb, const float* sub_weight,
                      const float* sub_weight_mom, void* stream) {
  if (int rc = require_sm100()) return rc;
  NvtxScope nvtx_scope("pfc_scatter_rows2");
  PFC_REQUIRE(weight && index && sub_weight && n_index >= 0 && emb > 0 && emb % 4 == 0, PFC_E_ARG, "pfc_scatter_rows2: bad argument");
  if (n_index == 0) return 0;
  move_rows2_kernel<false><<<row_grid(n_index), 256, 0, as_stream(stream)>>>(weight, weight_mom, index, n_index, emb,
                                                                              const_cast<float*>(sub_weight), const_cast<float*>(sub_weight_mom));
  PFC_LAUNCH_CHECK();
  return 0;
}

int pfc_cosface_dense(float* cosine, const int64_t* label, int64_t n_rows, int64_t n_classes, float s, float m, float* out, void* stream) {
  if (int rc = require_sm100()) return rc;
  PFC_REQUIRE(cosine && label && out && n_rows >= 0 && n_classes >= 0, PFC_E_ARG, "pfc_cosface_dense: bad argument");
  if (n_rows == 0 || n_classes == 0) return 0;
  cudaStream_t st = as_stream(stream);
  cosface_margin_kernel<<<(int)((n_rows + 255) / 256), 256, 0, st>>>(cosine, label, n_rows, n_classes, m);
  PFC_LAUNCH_CHECK();
  int64_t n = n_rows * n_classes;
  int64_t blocks = (n + 255) / 256;
  if (blocks > (int64_t)sm_count() * 16) blocks = (int64_t)sm_count() * 16;
  scale_kernel<<<(int)blocks, 256, 0, st>>>(cosine, s, n, out);
  PFC_LAUNCH_CHECK();
  return 0;
}

int pfc_arcface_dense(float* cosine, const int64_t* label, int64_t n_rows, int64_t n_classes, float s, float m, void* stream) {
  if (int rc = require_sm100()) return rc;
  PFC_REQUIRE(cosine && label && n_rows >= 0 && n_classes >= 0, PFC_E_ARG, "pfc_arcface_dense: bad argument");
  if (n_rows == 0 || n_classes == 0) return 0;
  int64_t n = n_rows * n_classes;
  int64_t blocks = (n + 255) / 256;
  if (blocks > (int64_t)sm_count() * 16) blocks = (int64_t)sm_count() * 16;
  arcface_dense_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(cosine, label, n_rows, n_classes, s, m);
  PFC_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
