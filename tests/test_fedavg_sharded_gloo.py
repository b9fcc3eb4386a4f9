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
