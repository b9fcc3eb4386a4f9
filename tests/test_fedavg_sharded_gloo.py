"""World-size-2 run of FedPavg_sharded's HOST logic on gloo/CPU: global weight normalisation, the flat partial-sum
buffer and the single all-reduce.  The per-rank weighted sum is supplied by the oracle (tests only); the result must
equal the oracle's sequential FedPavg (server.py:25-34) over all clients to fp32 rounding."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _models(k, seed=7):
    g = torch.Generator().manual_seed(seed)
    base = {"conv.weight": torch.randn(16, 3, 3, 3, generator=g), "bn.weight": torch.randn(16, generator=g),
            "bn.num_batches_tracked": torch.tensor(0, dtype=torch.int64), "fc.weight": torch.randn(37, 11, generator=g)}
    out = []
    for i in range(k):
        sd = {n: (v + 0.01 * torch.randn(v.shape, generator=g)) if v.dtype == torch.float32 else v + 3 * i + 1 for n, v in base.items()}
        out.append(sd)
    return out, [6000 + 37 * i for i in range(k)]


def _oracle_segments(srcs, weights_f32, device, flat=False):
    """CPU stand-in for fedavg.weighted_sum_segments: fp32 multiply then fp32 add, client order."""
    offsets, tot = [], 0
    for group in srcs:
        offsets.append(tot)
        tot += (group[0].numel() + 3) // 4 * 4
    flat_buf = torch.zeros(max(tot, 4), dtype=torch.float32)
    outs = []
    for s, group in enumerate(srcs):
        acc = None
        for w, t in zip(weights_f32, group):
            term = torch.tensor(w, dtype=torch.float32) * t.to(torch.float32)
            acc = term if acc is None else acc + term
        view = flat_buf[offsets[s]:offsets[s] + acc.numel()].view(acc.shape)
        view.copy_(acc)
        outs.append(view)
    return (outs, flat_buf) if flat else outs


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    try:
        from fedfr_b200.fedavg import FedPavg_sharded
        models, weights = _models(5)
        mine = list(range(rank, 5, world))                     # ragged: 3 clients on rank 0, 2 on rank 1
        out = FedPavg_sharded([models[i] for i in mine], [weights[i] for i in mine], device="cpu", _segments_fn=_oracle_segments)
        ret[rank] = {k: v.numpy().copy() for k, v in out.items()}
    finally:
        dist.barrier()
        dist.destroy_process_group()


def test_two_rank_sharded_fedavg_matches_sequential_reference():
    import __graft_entry__ as g
    g.build()
    sys.path.insert(0, ROOT)
    from oracle import partial_fc_oracle as O
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29751, ret), nprocs=2, join=True)
    models, weights = _models(5)
    ref = O.fedpavg([{k: v.numpy() for k, v in m.items()} for m in models], weights)
    for r in (0, 1):
        assert set(ret[r].keys()) == set(ref.keys())
        for k in ref:
            assert ret[r][k].dtype == np.float32 and ret[r][k].shape == np.asarray(ref[k]).shape
            np.testing.assert_allclose(ret[r][k], np.asarray(ref[k], dtype=np.float32), rtol=2e-6, atol=1e-6)
    for k in ref:                                              # every rank returns the same bits
        assert np.array_equal(ret[0][k], ret[1][k])


def test_output_views_are_cut_by_one_split():
    """Host logic of the output dict: one split_with_sizes over the flat buffer, gaps (16-byte padding) skipped, shapes kept."""
    sys.path.insert(0, ROOT)
    from fedfr_b200.fedavg import _split_plan, _split_views
    buf = torch.arange(64.)
    spans = [(0, 3, (3,)), (4, 4, (2, 2)), (8, 0, (0,)), (8, 1, ()), (12, 5, (5,)), (20, 2, (2, 1)), (40, 24, (2, 3, 4))]
    plan = _split_plan(spans, buf.numel())
    assert sum(plan[0]) == buf.numel()
    views = _split_views(buf, plan)
    for (off, n, shape), v in zip(spans, views):
        assert tuple(v.shape) == tuple(shape)
        assert torch.equal(v.reshape(-1), buf[off:off + n])
        assert n == 0 or v.data_ptr() == buf.data_ptr() + 4 * off
    # spans may come in any order (state_dict key order interleaves the fp32 and the int64 region)
    shuffled = [spans[i] for i in (3, 0, 6, 1, 5, 4, 2)]
    for (off, n, shape), v in zip(shuffled, _split_views(buf, _split_plan(shuffled, buf.numel()))):
        assert tuple(v.shape) == tuple(shape) and torch.equal(v.reshape(-1), buf[off:off + n])
    try:
        _split_plan([(0, 4, (4,)), (2, 4, (4,))], 8)
    except ValueError:
        pass
    else:
        raise AssertionError("overlapping spans must be rejected")


def test_weight_rounding_matches_python_float_times_fp32_tensor():
    """server.py:31 multiplies a Python double with an fp32 tensor: torch rounds the scalar to fp32 first."""
    sys.path.insert(0, ROOT)
    from fedfr_b200.fedavg import _as_f32
    rng = np.random.default_rng(3)
    ws = list(rng.random(2000) * 10.0 ** rng.integers(-30, 30, 2000)) + [0.1, 1 / 3, 6000 / 271590.0, 1e-46, 3e38]
    got = _as_f32([float(w) for w in ws])
    for w, g in zip(ws, got):
        t = torch.ones(1) * float(w)                      # the reference's own arithmetic
        assert float(t) == g, (w, g, float(t))


def test_flat_state_dict_host_path_with_a_stand_in_kernel(monkeypatch):
    """The FlatStateDict call end to end on the CPU: layout, pointer tables, padding between the fp32 and the int64 region,
    the output views -- with ``_launch`` replaced by a numpy loop that reads the very tables the CUDA kernel would read
    (fp32 multiply, fp32 add, client order, 0 + first term: server.py:25-34)."""
    import ctypes as C
    sys.path.insert(0, ROOT)
    from fedfr_b200 import fedavg as FA
    from fedfr_b200 import _native as N
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)          # no driver in the CPU container

    def stand_in(n_seg, k, device):
        t = FA._tables
        for s in range(n_seg):
            n = int(t.len_np[s])
            out = np.ctypeslib.as_array((C.c_float * n).from_address(int(t.out_np[s])))
            acc = np.zeros(n, dtype=np.float32)
            for i in range(k):
                ptr = int(t.src_np[s * k + i])
                if int(t.dtype_np[s]) & 0xff == N.FEDAVG_F32:
                    v = np.ctypeslib.as_array((C.c_float * n).from_address(ptr))
                else:
                    v = np.ctypeslib.as_array((C.c_int64 * n).from_address(ptr)).astype(np.float32)
                acc = acc + np.float32(t.w_np[i]) * v
            out[:] = acc

    monkeypatch.setattr(FA, "_launch", stand_in)
    models, weights = _models(5)
    models = [dict(m, **{"odd.bias": torch.randn(7), "odd2.weight": torch.randn(3, 5)}) for m in models]     # sizes that need padding
    flats = [FA.flatten_state_dict(m) for m in models]
    assert all(f.layout is flats[0].layout for f in flats)
    wn = FA._as_f32(FA._normalised_weights(weights))
    out, flat_buf = FA._weighted_sum_flat(flats, wn, torch.device("cpu"))
    assert list(out.keys()) == list(models[0].keys())
    for name in models[0]:
        ref = 0
        for w, m in zip([w / sum(weights) for w in weights], models):
            ref = ref + w * m[name]                                   # the reference's own expression (server.py:31-32)
        assert out[name].dtype == torch.float32 and out[name].shape == models[0][name].shape
        assert torch.equal(out[name], ref.to(torch.float32)), name
        lo, hi = flat_buf.data_ptr(), flat_buf.data_ptr() + 4 * flat_buf.numel()
        assert lo <= out[name].data_ptr() < hi and (models[0][name].dtype == torch.int64 or out[name].data_ptr() % 16 == 0)
