"""Hard-negative mining by similarity threshold (client.py:208-215, :232-235).

Golden: tests/golden/hardneg.npz -- ``HN_ID`` selected by the UNMODIFIED ``Client.choose_hard_negative`` on CPU, and the
index list of the feature-based expression (tests/golden/make_golden.py hardneg).  A float32 product may fall on either
side of the threshold within rounding, so comparisons allow exactly the columns the float64 oracle marks ambiguous
(|best cosine - threshold| <= 2e-6); everything else must agree exactly."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__ as g
    g.build()
    import fedfr_b200
    return fedfr_b200


@pytest.fixture(scope="module")
def golden():
    z = np.load(os.path.join(HERE, "golden", "hardneg.npz"))
    return {k: z[k] for k in z.files}


def _agrees(got, certain, ambiguous):
    got, allowed = set(got.tolist()), set(certain.tolist()) | set(ambiguous.tolist())
    return set(certain.tolist()) <= got <= allowed


def test_oracle_matches_reference_golden(golden):
    from oracle import hardneg_oracle as O
    c, a = O.similar_columns(O.normalize(golden["self_fc"]), O.normalize(golden["pretrain_fc"]), float(golden["threshold"]))
    assert _agrees(golden["HN_ID"], c, a) and len(a) <= 2
    c, a = O.similar_columns(golden["local_feats"], golden["pretrained_feats"], float(golden["threshold2"]))
    assert _agrees(golden["unique_idx"], c, a) and len(a) <= 2
    assert O.similar_columns(np.zeros((0, 4)), np.zeros((3, 4)), 0.2)[0].size == 0


def test_host_logic_cpu(pkg, golden):
    from oracle import hardneg_oracle as O
    ids = pkg.hard_negative_ids(torch.from_numpy(golden["self_fc"]), torch.from_numpy(golden["pretrain_fc"]), 0.2, _mask_fn=O.mask_fn)
    assert ids.dtype == np.int64 and np.array_equal(ids, golden["HN_ID"])            # same float32 matmul as the reference
    idx = pkg.similar_columns(torch.from_numpy(golden["local_feats"]), torch.from_numpy(golden["pretrained_feats"]),
                              float(golden["threshold2"]), _mask_fn=O.mask_fn)
    assert np.array_equal(idx, golden["unique_idx"]) and np.all(np.diff(idx) > 0)
    # what the reference derives from the ids (client.py:249-255): 1-based image list of the selected identities
    assert np.array_equal(np.nonzero(np.isin(golden["pretrain_label"], ids))[0] + 1, golden["imgidx"])
    with pytest.raises(NotImplementedError):
        pkg.hard_negative_ids(torch.zeros(2, 4), torch.zeros(3, 4), 5)
    with pytest.raises(RuntimeError):
        pkg.similar_columns(torch.zeros(2, 4), torch.zeros(3, 4), 0.2)               # no CPU fallback


@pytest.mark.gpu
def test_gpu_matches_reference_golden(pkg, golden):
    from oracle import hardneg_oracle as O
    dev = "cuda:0"
    ids = pkg.hard_negative_ids(torch.from_numpy(golden["self_fc"]).to(dev), torch.from_numpy(golden["pretrain_fc"]).to(dev), 0.2)
    c, a = O.similar_columns(O.normalize(golden["self_fc"]), O.normalize(golden["pretrain_fc"]), 0.2, band=1e-5)
    assert _agrees(ids, c, a) and len(set(ids.tolist()) ^ set(golden["HN_ID"].tolist())) <= len(a)
    idx = pkg.similar_columns(torch.from_numpy(golden["local_feats"]).to(dev), torch.from_numpy(golden["pretrained_feats"]).to(dev),
                              float(golden["threshold2"]))
    c, a = O.similar_columns(golden["local_feats"], golden["pretrained_feats"], float(golden["threshold2"]), band=1e-5)
    assert _agrees(idx, c, a) and np.all(np.diff(idx) > 0)


@pytest.mark.gpu
@pytest.mark.parametrize("na,nb,emb,thr", [(300, 5000, 512, 0.2), (65, 129, 33, 0.0), (1, 64, 512, -1.5), (200, 70, 16, 2.0), (0, 10, 8, 0.1)])
def test_gpu_matches_oracle(pkg, na, nb, emb, thr):
    from oracle import hardneg_oracle as O
    g = torch.Generator().manual_seed(na + nb)
    a = torch.nn.functional.normalize(torch.randn(na, emb, generator=g))
    b = torch.nn.functional.normalize(torch.randn(nb, emb, generator=g) + (0.21 * emb ** 0.5 * a[torch.randint(0, max(na, 1), (nb,), generator=g)] if na else 0))
    got = pkg.similar_columns(a.to("cuda:0"), b.to("cuda:0"), thr)
    c, amb = O.similar_columns(a.numpy(), b.numpy(), thr, band=1e-5)      # 512-step fp32 FMA chain: gamma_512 * sum|ab| < 1e-5 here
    assert _agrees(got, c, amb) and len(amb) <= 5
    if (na, nb) == (300, 5000):
        assert 0.2 * nb < len(got) < 0.8 * nb                 # a real mix of hits and misses
    if thr == -1.5 and na:
        assert len(got) == nb
    if thr == 2.0 or na == 0:
        assert len(got) == 0
