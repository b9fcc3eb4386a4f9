"""Pairwise-cosine ROC histogram (roc_cuda.py:14-28): oracle vs the reference golden, host logic on CPU / gloo, and the
sm_100a kernel through the C ABI against both (integer-exact).  Golden: tests/golden/roc.npz, produced by the unmodified
reference kernel (tests/golden/make_golden_roc.py)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CASES = ["a", "b", "c"]


@pytest.fixture(scope="module")
def R():
    import __graft_entry__ as g
    g.build()
    from oracle import roc_oracle
    return roc_oracle


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "roc.npz"))


def _case(z, c):
    return z[c + "/feature"], z[c + "/label"], int(z[c + "/batch_size"]), int(z[c + "/target_size"])


def _synthetic(n, emb, n_ids, seed):
    rng = np.random.default_rng(seed)
    centres = rng.standard_normal((n_ids, emb))
    label = np.sort(rng.integers(0, n_ids, n)).astype(np.int32)
    f = centres[label] + 0.9 * rng.standard_normal((n, emb))
    return (f / np.linalg.norm(f, axis=1, keepdims=True)).astype(np.float32), label


# ----------------------------------------------------------------------------------------------- oracle (CPU)

@pytest.mark.parametrize("c", CASES)
def test_oracle_matches_reference_kernel(R, golden, c):
    f, l, bs, t = _case(golden, c)
    assert np.array_equal(R.roc_histogram_batched(f, l, bs, t), golden[c + "/hist"])
    for b, start in enumerate(range(0, t, bs)):                 # each launch of the reference's consumer loop
        stop = min(start + bs, t)
        assert np.array_equal(R.roc_histogram(f[start:], l[start:], f[start:stop], l[start:stop]), golden[c + "/per_batch"][b])


@pytest.mark.parametrize("c", CASES)
def test_batch_loop_is_one_call_and_shards_add_up(R, golden, c):
    f, l, _, t = _case(golden, c)
    whole = R.roc_histogram(f, l, f[:t], l[:t])
    assert np.array_equal(whole, golden[c + "/hist"])
    cut = t // 3
    parts = R.roc_histogram(f, l, f[:cut], l[:cut], 0) + R.roc_histogram(f, l, f[cut:t], l[cut:t], cut)
    assert np.array_equal(parts, whole)


def test_oracle_empty_and_out_of_range(R):
    f, l = _synthetic(5, 8, 2, 0)
    assert R.roc_histogram(f, l, f[:0], l[:0]).sum() == 0
    assert R.roc_histogram(f[:0], l[:0], f[:0], l[:0]).sum() == 0
    with pytest.raises(ValueError):
        R.roc_histogram(2 * f, l, 2 * f, l)


# ----------------------------------------------------------------------------------------------- host logic (CPU)

def test_tpr_at_fpr_matches_plot_roc(golden):
    from fedfr_b200.roc import tpr_at_fpr
    assert tpr_at_fpr(golden["a/hist"]) == [float(x) for x in golden["a/tpr"]]
    assert tpr_at_fpr(golden["a/hist"].reshape(-1, 2)) == [float(x) for x in golden["a/tpr"]]


def test_shard_rows_cover_and_balance():
    from fedfr_b200.roc import shard_rows
    for t, n, w in [(48, 80, 2), (1, 33, 2), (70, 70, 4), (0, 10, 2), (800, 100000, 8), (5, 5, 8)]:
        sh = shard_rows(t, n, w)
        assert len(sh) == w and sh[0][0] == 0 and sh[-1][1] == t
        assert all(a[1] == b[0] for a, b in zip(sh, sh[1:])) and all(r0 <= r1 for r0, r1 in sh)
    sh = shard_rows(4000, 4000, 4)
    pairs = [sum(4000 - 1 - i for i in range(r0, r1)) for r0, r1 in sh]
    assert max(pairs) - min(pairs) <= 2 * 4000                 # within one row's worth of pairs of each other


def test_cpu_tensors_are_refused():
    from fedfr_b200.roc import calc_ROC
    f = torch.zeros(4, 8)
    l = torch.zeros(4, dtype=torch.int32)
    with pytest.raises(RuntimeError):
        calc_ROC(f, l, f, l, torch.zeros(4002, dtype=torch.int64))


def _oracle_calc(feature, label, subfeature, sublabel, out, sub_offset=0):
    """CPU stand-in for fedfr_b200.roc.calc_ROC in the gloo test (tests only)."""
    from oracle import roc_oracle
    h = roc_oracle.roc_histogram(feature.numpy(), label.numpy(), subfeature.numpy(), sublabel.numpy(), sub_offset)
    out += torch.from_numpy(h)
    return out


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fedfr_b200.roc import roc_histogram
        z = np.load(os.path.join(HERE, "golden", "roc.npz"))
        f, l, _, t = _case(z, "a")
        ret[rank] = roc_histogram(torch.from_numpy(f), torch.from_numpy(l), target_size=t, _calc=_oracle_calc)
    finally:
        dist.barrier()
        dist.destroy_process_group()


def test_two_rank_histogram_host_logic(R, golden):
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, 29771, ret), nprocs=2, join=True)
    for r in (0, 1):
        assert ret[r].shape == (2001, 2) and ret[r].dtype == np.int64
        assert np.array_equal(ret[r].reshape(-1), golden["a/hist"])


# ----------------------------------------------------------------------------------------------- CUDA kernel (C ABI)

def _dev(a, dtype):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda:0", dtype)


@pytest.mark.gpu
@pytest.mark.parametrize("c", CASES)
def test_kernel_matches_reference_golden(golden, c):
    from fedfr_b200.roc import calc_ROC, roc_histogram
    f, l, bs, t = _case(golden, c)
    assert np.array_equal(roc_histogram(f, l, target_size=t, batch_size=bs).reshape(-1), golden[c + "/hist"])
    fd, ld = _dev(f, torch.float32), _dev(l, torch.int32)
    for b, start in enumerate(range(0, t, bs)):                 # the reference's launches, one by one
        stop = min(start + bs, t)
        out = torch.zeros(4002, dtype=torch.int64, device="cuda:0")
        calc_ROC(fd[start:], ld[start:], fd[start:stop], ld[start:stop], out)
        assert np.array_equal(out.cpu().numpy(), golden[c + "/per_batch"][b])


@pytest.mark.gpu
@pytest.mark.parametrize("n,emb,t,seed", [(700, 512, 300, 1), (131, 33, 131, 2), (65, 1, 64, 3), (200, 512, 1, 4)])
def test_kernel_matches_oracle(R, n, emb, t, seed):
    from fedfr_b200.roc import calc_ROC, roc_histogram
    f, l = _synthetic(n, emb, 7, seed)
    want = R.roc_histogram(f, l, f[:t], l[:t])
    assert np.array_equal(roc_histogram(f, l, target_size=t).reshape(-1), want)
    fd, ld = _dev(f, torch.float32), _dev(l, torch.int32)        # out is accumulated into; shards add up
    out = torch.zeros(4002, dtype=torch.int64, device="cuda:0")
    cut = t // 2
    calc_ROC(fd, ld, fd[:cut], ld[:cut], out, sub_offset=0)
    calc_ROC(fd, ld, fd[cut:t], ld[cut:t], out, sub_offset=cut)
    assert np.array_equal(out.cpu().numpy(), want)


@pytest.mark.gpu
def test_kernel_full_size_properties():
    """At an evaluation-sized input the oracle is too slow: check what must hold for any input."""
    from fedfr_b200.roc import roc_histogram
    n, t = 20000, 4000
    f, l = _synthetic(n, 512, 400, 5)
    h = roc_histogram(f, l, target_size=t, batch_size=800)
    assert h.sum() == t * (t - 1) // 2 + t * (n - t)
    li = l.astype(np.int64)                                      # same-ID pairs (i < t, i < j), counted from the labels
    after = np.zeros(n, dtype=np.int64)
    seen = {}
    for j in range(n - 1, -1, -1):
        after[j] = seen.get(li[j], 0)
        seen[li[j]] = after[j] + 1
    assert h[:, 0].sum() == after[:t].sum()
    again = roc_histogram(f, l, target_size=t)
    assert np.array_equal(h, again)                              # atomics only reorder integer adds


@pytest.mark.gpu
def test_kernel_empty_inputs_and_bad_types():
    from fedfr_b200.roc import calc_ROC
    f = torch.zeros(0, 16, device="cuda:0")
    l = torch.zeros(0, dtype=torch.int32, device="cuda:0")
    out = torch.zeros(4002, dtype=torch.int64, device="cuda:0")
    assert int(calc_ROC(f, l, f, l, out).sum()) == 0
    g = torch.ones(3, 16, device="cuda:0") / 4
    gl = torch.zeros(3, dtype=torch.int32, device="cuda:0")
    assert int(calc_ROC(g, gl, f, l, out).sum()) == 0
    with pytest.raises(TypeError):
        calc_ROC(g.double(), gl, g.double(), gl, out)
    with pytest.raises(TypeError):
        calc_ROC(g, gl.long(), g, gl.long(), out)
    calc_ROC(g, gl, g, gl, out)                                   # three identical unit rows: 3 pairs at cosine 1.0
    assert int(out[2 * 2000]) == 3 and int(out.sum()) == 3
