// Warp-level row normalisation shared by the stand-alone kernel (rows.cu) and the normaliser warps that ride inside
// the forward logits kernel (tc_kernels.cu).        normalize(sub_weight): partial_fc.py:127 (+ gather :105)
#pragma once
#include "common.cuh"

namespace pfc {

// Warp `warp` of `n_warps` normalises rows warp*kRows + [0, kRows), then strides by n_warps*kRows.
// kRows rows are in flight together (kRows * kVecPerLane 128-bit loads per lane before the first use).
// Algorithmic bytes per row: read 4E, write 2E (bf16) + 4.
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {   // bytes: multiple of 16
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// kPrefetch > 0: the rows of iteration (it + kPrefetch) are pulled into L2 by a bulk prefetch (no registers, no shared
// memory), so that a few warps per SM keep enough HBM requests in flight (used by the normaliser warps, which have
// only eight warps per SM to hide the DRAM latency with).
template <int kVecPerLane, int kRows, int kPrefetch = 0>
__device__ __forceinline__ void normalize_rows_warp(const float* __restrict__ w, const int64_t* __restrict__ index, int64_t n_rows, int emb,
                                                    __nv_bfloat16* __restrict__ out_bf16, float* __restrict__ out_f32,
                                                    float* __restrict__ inv_norm, int64_t warp, int64_t n_warps, int lane) {
  const int nvec = emb >> 2;
  if (kPrefetch > 0) {
#pragma unroll 1
    for (int a = 0; a < kPrefetch; ++a) {
      const int64_t r = (warp + a * n_warps) * kRows + lane;
      if (lane < kRows && r < n_rows) prefetch_l2_bulk(w + (index ? index[r] : r) * emb, (uint32_t)emb * 4u);
    }
  }
  for (int64_t r0 = warp * kRows; r0 < n_rows; r0 += n_warps * kRows) {
    if (kPrefetch > 0) {
      const int64_t r = r0 + (int64_t)kPrefetch * n_warps * kRows + lane;
      if (lane < kRows && r < n_rows) prefetch_l2_bulk(w + (index ? index[r] : r) * emb, (uint32_t)emb * 4u);
    }
    float4 v[kRows][kVecPerLane];
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
      const int64_t r = r0 + k;
      if (r < n_rows) {
        const int64_t src = index ? index[r] : r;
        const float4* p = reinterpret_cast<const float4*>(w + src * emb);
#pragma unroll
        for (int i = 0; i < kVecPerLane; ++i) {
          const int c = lane + i * 32;
          if (c < nvec) v[k][i] = ld_stream_f4(p + c);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
      const int64_t r = r0 + k;
      if (r < n_rows) {
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < kVecPerLane; ++i) {
          const int c = lane + i * 32;
          if (c < nvec) ss += v[k][i].x * v[k][i].x + v[k][i].y * v[k][i].y + v[k][i].z * v[k][i].z + v[k][i].w * v[k][i].w;
        }
        ss = warp_sum(ss);
        const float nrm = fmaxf(sqrtf(ss), 1e-12f);
        const float inv = 1.0f / nrm;
        if (lane == 0 && inv_norm) inv_norm[r] = inv;
#pragma unroll
        for (int i = 0; i < kVecPerLane; ++i) {
          const int c = lane + i * 32;
          if (c < nvec) {
            // fp32 output (check mode): divide, so it matches F.normalize bit for bit.  bf16-only output: multiply by the
            // reciprocal (1 ulp of fp32, far below the bf16 rounding) -- the division sequence costs ~10 issue slots each.
            float4 o = out_f32 ? make_float4(v[k][i].x / nrm, v[k][i].y / nrm, v[k][i].z / nrm, v[k][i].w / nrm)
                               : make_float4(v[k][i].x * inv, v[k][i].y * inv, v[k][i].z * inv, v[k][i].w * inv);
            if (out_bf16) {
              uint2 pk = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
              *reinterpret_cast<uint2*>(out_bf16 + r * emb + c * 4) = pk;
            }
            if (out_f32) st_stream_f4(reinterpret_cast<float4*>(out_f32 + r * emb) + c, o);
          }
        }
      }
    }
  }
}

}  // namespace pfc
