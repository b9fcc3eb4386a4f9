"""2-rank NCCL run of the CUDA path against the goldens the unmodified reference produced on 2 gloo ranks
(class sharding with a ragged split, gathered labels/features, stats exchange, reduce-scatter, x W).
Needs >= 2 GPUs; skipped otherwise."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-7 * b.size ** 0.5))


def _worker(rank, world, name, check_mode, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import fedfr_b200
        from golden_util import Case, margin_of
        case = Case(name)
        cfg = case.cfg
        dev = torch.device("cuda", rank)
        head = fedfr_b200.PartialFC(rank, rank, world, cfg["batch"], False, margin_of(fedfr_b200, cfg), cfg["num_classes"],
                                    sample_rate=cfg["sample_rate"], embedding_size=cfg["emb"], prefix="/tmp", check_mode=check_mode)
        head.weight.copy_(case.weights[rank].to(dev))
        head.weight_mom.zero_()
        opt = torch.optim.SGD([{"params": head.parameters()}], lr=cfg["lr"], momentum=0.9, weight_decay=5e-4)
        errs = {}
        real_rand = torch.rand
        for step in range(cfg["steps"]):
            if case.has(rank, step, "perm"):
                perm = torch.from_numpy(case.get(rank, step, "perm")).to(dev)
                torch.rand = lambda *a, **k: perm.clone()
            x_grad, loss = head.forward_backward(case.labels[rank].to(dev), case.features[rank].to(dev), opt)
            torch.rand = real_rand
            if case.has(rank, step, "index"):
                assert np.array_equal(head.index.cpu().numpy(), case.get(rank, step, "index"))
            errs[f"loss{step}"] = abs(float(loss) - float(case.get(rank, step, "loss"))) / max(1.0, abs(float(case.get(rank, step, "loss"))))
            errs[f"dx{step}"] = _rel(x_grad.cpu().numpy(), case.get(rank, step, "x_grad"))
            errs[f"dw{step}"] = _rel(head.sub_weight.grad.cpu().numpy(), case.get(rank, step, "dw"))
            opt.step()
            head.update()
            opt.zero_grad()
            head.weight.copy_(torch.from_numpy(case.get(rank, step, "weight_after")).to(dev))
            head.weight_mom.copy_(torch.from_numpy(case.get(rank, step, "mom_after")).to(dev))
        ret[rank] = errs
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("check_mode", [True, False])
@pytest.mark.parametrize("name,port", [("w2_sr1_ragged", 29741), ("w2_sr03", 29742), ("w2_arc_sr03", 29743)])
def test_two_gpu_matches_reference(name, port, check_mode):
    import __graft_entry__ as g
    g.build()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, name, check_mode, port + (10 if check_mode else 0), ret), nprocs=2, join=True)
    tol = 1e-4 if check_mode else 1e-2
    for r in (0, 1):
        for k, v in ret[r].items():
            assert v < tol, (r, k, v, dict(ret[r]))


def _fedavg_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import fedfr_b200
        dev = torch.device("cuda", rank)
        g = torch.Generator().manual_seed(11)
        K = 6
        base = {"a.weight": torch.randn(300, 77, generator=g), "bn.num_batches_tracked": torch.tensor(5, dtype=torch.int64),
                "b.bias": torch.randn(1003, generator=g)}
        models = [{n: (v + 0.01 * torch.randn(v.shape, generator=g)) if v.dtype == torch.float32 else v + i for n, v in base.items()}
                  for i in range(K)]
        weights = [100 + 7 * i for i in range(K)]
        mine = list(range(rank, K, world))
        out = fedfr_b200.FedPavg_sharded([{n: v.to(dev) for n, v in models[i].items()} for i in mine], [weights[i] for i in mine])
        full = fedfr_b200.FedPavg([{n: v.to(dev) for n, v in m.items()} for m in models], weights)      # sequential order, one GPU
        ret[rank] = max(float((out[n] - full[n]).abs().max() / full[n].abs().max().clamp_min(1e-20)) for n in full)
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_sharded_fedavg():
    import __graft_entry__ as g
    g.build()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_fedavg_worker, args=(2, 29761, ret), nprocs=2, join=True)
    assert max(ret.values()) < 5e-6, dict(ret)


def _roc_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from fedfr_b200.roc import roc_histogram
        z = np.load(os.path.join(HERE, "golden", "roc.npz"))
        out = {}
        for c in "abc":          # each rank takes a row range of the sub block (by pair count), one int64 all-reduce
            out[c] = roc_histogram(z[c + "/feature"], z[c + "/label"], target_size=int(z[c + "/target_size"]),
                                   device=torch.device("cuda", rank))
        ret[rank] = out
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_roc_histogram():
    """roc_cuda.py's multi-GPU fan-out (one worker process per GPU, roc_cuda.py:89-108): both ranks return the reference total."""
    import __graft_entry__ as g
    g.build()
    ret = mp.Manager().dict()
    mp.spawn(_roc_worker, args=(2, 29766, ret), nprocs=2, join=True)
    z = np.load(os.path.join(HERE, "golden", "roc.npz"))
    for r in (0, 1):
        for c in "abc":
            assert np.array_equal(ret[r][c].reshape(-1), z[c + "/hist"]), (r, c)


def _selfcheck_worker(rank, world, port, sr, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import fedfr_b200
        from fedfr_b200 import selfcheck as SC
        dev = torch.device("cuda", rank)
        B, C, E = 512, 200_001, 512                       # ragged shards, the kernels' production tile shapes
        torch.manual_seed(100 + rank)
        head = fedfr_b200.PartialFC(rank, rank, world, B, False, fedfr_b200.CosFace(64.0, 0.4), C, sample_rate=sr, embedding_size=E, prefix="/tmp")
        x = torch.nn.functional.normalize(torch.randn(B, E, device=dev))
        y = torch.randint(0, C, (B,), device=dev)
        opt = torch.optim.SGD([{"params": head.parameters()}], lr=0.1, momentum=0.9)
        for _ in range(2):
            head.sub_weight.grad = None
            xg, loss = head.forward_backward(y, x, opt)
        out = SC.check_head_step(head, y, x, xg, loss)
        ret[rank] = out
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("sr,port", [(1.0, 29771), (0.1, 29772)])
def test_two_gpu_selfcheck_at_production_tiles(sr, port):
    """The check bench.py runs after its timed loops (fedfr_b200/selfcheck.py: reference arithmetic at 1e-2, bf16-emulating
    restatement row by row, identical loss bits on every rank) on 2 NCCL ranks, unsampled and sampled."""
    import __graft_entry__ as g
    g.build()
    ret = mp.Manager().dict()
    mp.spawn(_selfcheck_worker, args=(2, port, sr, ret), nprocs=2, join=True)
    for r in (0, 1):
        assert ret[r]["ok"] and ret[r]["ok_all_ranks"] and ret[r]["loss_identical_on_all_ranks"], dict(ret[r])
