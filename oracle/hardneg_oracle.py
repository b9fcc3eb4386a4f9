"""CPU oracle for hard-negative mining by similarity threshold  --  TEST INFRASTRUCTURE ONLY (see partial_fc_oracle.py).

Restates ``similarity = matmul(a, b.t()); unique(torch.where(similarity > threshold)[1])`` (client.py:208-215, :232-235)
in float64 and reports which columns are decided beyond rounding: a float32 product (the reference's CPU sgemm, or the
CUDA kernel's FMA chain) may land on either side of the threshold only for columns whose best cosine is within ``band``
of it.  Pinned by ``tests/golden/hardneg.npz`` (the unmodified ``Client.choose_hard_negative`` run on CPU with stub
loader/logger objects, ``tests/golden/make_golden.py hardneg``)."""
import numpy as np


def similar_columns(a, b, threshold, band=2e-6):
    """-> (certain, ambiguous): sorted int64 column ids with max_i <a_i, b_j> > threshold + band, and those within band."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.shape[0] == 0 or b.shape[0] == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    best = np.full(b.shape[0], -np.inf)
    for s in range(0, a.shape[0], 256):                     # bounded temporaries
        best = np.maximum(best, (a[s:s + 256] @ b.T).max(axis=0))
    certain = np.nonzero(best > threshold + band)[0].astype(np.int64)
    ambiguous = np.nonzero(np.abs(best - threshold) <= band)[0].astype(np.int64)
    return certain, ambiguous


def normalize(x, eps=1e-12):
    x = np.asarray(x, dtype=np.float64)
    return x / np.maximum(np.linalg.norm(x, axis=1, keepdims=True), eps)


def mask_fn(a, b, threshold):
    """CPU stand-in for fedfr_b200.hardneg._hit_mask (host-logic tests only): float32 matmul like the reference."""
    import torch
    return (torch.matmul(a.float(), b.float().t()) > threshold).any(dim=0).to(torch.uint8)
