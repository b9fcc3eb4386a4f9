/* CPU oracle (plain C) for the pairwise-cosine ROC histogram  --  TEST INFRASTRUCTURE ONLY.
 *
 * The checker, never the product: only tests/, __graft_entry__.smoke() and bench-style tools may load the library
 * built from this file (oracle/_build/liboracle_c.so, compiled by __graft_entry__.build() with -ffp-contract=off).
 *
 * Restates the arithmetic of the reference's numba kernel calc_ROC (roc_cuda.py:14-28):
 *   i, j = cuda.grid(2); guard  i < j, j < feature.shape[0], i < sublabel.shape[0]          roc_cuda.py:16-17
 *   tmp = 0. (a double); tmp += subfeature[i,k] * feature[j,k] (float32 product)             roc_cuda.py:18-20
 *   index_dis = int((tmp + 1) * 1000)                                                        roc_cuda.py:21
 *   out[2*index_dis] += 1 when the labels agree, else out[2*index_dis + 1] += 1              roc_cuda.py:24-28
 * `sub_offset` generalises the guard to (sub_offset + i) < j (0 = the reference); see include/fedfr_b200.h.
 * Pinning: tests/test_oracle_golden.py::test_roc_* checks it against tests/golden/roc.npz, produced by running the
 * unmodified reference kernel under numba's CUDA simulator (tests/golden/make_golden_roc.py). */
#include <stdint.h>

int oracle_roc_histogram(const float* feature, const int32_t* label, int64_t n, const float* subfeature,
                         const int32_t* sublabel, int64_t n_sub, int64_t sub_offset, int emb, int64_t* hist) {
  for (int64_t i = 0; i < n_sub; ++i) {
    const float* a = subfeature + i * (int64_t)emb;
    for (int64_t j = sub_offset + i + 1; j < n; ++j) {
      const float* b = feature + j * (int64_t)emb;
      double tmp = 0.0;
      for (int k = 0; k < emb; ++k) {
        const float p = a[k] * b[k];
        tmp += (double)p;
      }
      int bin = (int)((tmp + 1.0) * 1000.0);
      if (bin < 0 || bin > 2000) return -1;          /* the reference would write out of bounds */
      hist[2 * bin + (sublabel[i] != label[j] ? 1 : 0)] += 1;
    }
  }
  return 0;
}
