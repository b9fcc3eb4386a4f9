"""The two-tier ROC kernel (the default since round 2, pfc_set_roc_mode(1): fp32 FMA filter + exact chain near bin edges,
csrc/roc.cu) must return the same integers as the exact kernel (mode 0) and as the reference golden -- on the GPU.  (Its CPU twin is
tests/test_kernel_emulation.py::test_two_tier_roc_kernel_is_integer_identical.)"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
pytestmark = pytest.mark.gpu


@pytest.fixture()
def two_tier():
    import __graft_entry__ as g
    g.build()
    from fedfr_b200 import _native as N

    def set_mode(m):
        N.check(N.lib.pfc_set_roc_mode(m), "pfc_set_roc_mode")
    yield set_mode
    set_mode(1)


def test_two_tier_matches_golden_and_exact_kernel(two_tier):
    from fedfr_b200.roc import roc_histogram
    z = np.load(os.path.join(HERE, "golden", "roc.npz"))
    two_tier(1)
    for c in "abc":
        got = roc_histogram(z[c + "/feature"], z[c + "/label"], target_size=int(z[c + "/target_size"]))
        assert np.array_equal(got.reshape(-1), z[c + "/hist"]), c
    rng = np.random.default_rng(0)
    n, emb = 3000, 512
    f = rng.standard_normal((n, emb)).astype(np.float32)
    f /= np.linalg.norm(f, axis=1, keepdims=True)
    for r in range(0, 600, 3):                                  # rows whose cosines are exactly 0 / 1 / -1: on the bin edges
        f[r] = 0
        f[r, (r // 3) % emb] = 1.0
    f[1], f[7], f[10] = f[4], -f[4], 0.0
    f[2000:] *= rng.uniform(0.2, 1.0, (n - 2000, 1)).astype(np.float32)
    l = rng.integers(0, 50, n).astype(np.int32)
    fast = roc_histogram(f, l, target_size=1500)
    two_tier(0)
    exact = roc_histogram(f, l, target_size=1500)
    assert np.array_equal(fast, exact) and fast.sum() == 1500 * 1499 // 2 + 1500 * 1500
