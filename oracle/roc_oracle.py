"""ctypes face of ``oracle/oracle_c.c`` (ROC histogram)  --  TEST INFRASTRUCTURE ONLY, like everything under oracle/.

``roc_histogram`` restates ``roc_cuda.calc_ROC`` (roc_cuda.py:14-28) for one (feature, subfeature) block;
``roc_histogram_batched`` restates the producer/consumer loop around it (roc_cuda.py:30-53 and :136-139).
Pinned by ``tests/golden/roc.npz`` (the unmodified reference kernel under numba's CUDA simulator).
"""
import ctypes as C
import os

import numpy as np

_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "liboracle_c.so")
_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise ImportError(f"{_SO} missing: run __graft_entry__.build() (gcc) first")
        _lib = C.CDLL(_SO)
        _lib.oracle_roc_histogram.restype = C.c_int
        _lib.oracle_roc_histogram.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                              C.c_int64, C.c_int, C.c_void_p]
    return _lib


def roc_histogram(feature, label, subfeature, sublabel, sub_offset=0, out=None):
    feature = np.ascontiguousarray(feature, dtype=np.float32)
    subfeature = np.ascontiguousarray(subfeature, dtype=np.float32)
    label = np.ascontiguousarray(label, dtype=np.int32)
    sublabel = np.ascontiguousarray(sublabel, dtype=np.int32)
    if out is None:
        out = np.zeros(2001 * 2, dtype=np.int64)
    emb = feature.shape[1] if feature.ndim == 2 else 0
    rc = _load().oracle_roc_histogram(feature.ctypes.data, label.ctypes.data, feature.shape[0], subfeature.ctypes.data,
                                      sublabel.ctypes.data, subfeature.shape[0], int(sub_offset), emb, out.ctypes.data)
    if rc != 0:
        raise ValueError("cosine outside [-1, 1]: the reference kernel would write out of bounds")
    return out


def roc_histogram_batched(feature, label, batch_size, target_size):
    """Sum over the reference's batches: block b compares feature[b*bs : min((b+1)*bs, target)] with feature[b*bs:]."""
    out = np.zeros(2001 * 2, dtype=np.int64)
    for start in range(0, int(target_size), int(batch_size)):
        stop = min(start + int(batch_size), int(target_size))
        roc_histogram(feature[start:], label[start:], feature[start:stop], label[start:stop], 0, out)
    return out
