"""The device-side self-check bench.py runs after its timed loops (fedfr_b200/selfcheck.py) against the CPU oracle:
its fp32 flavour must equal ``oracle.forward_backward`` (the pinned restatement of the reference), its bf16-emulating
flavour ``oracle.forward_backward_bf16``; a 2-rank gloo run exercises the collectives and the failure path."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("margin,s,m", [("cosface", 64.0, 0.4), ("arcface", 30.0, 0.5)])
def test_reference_step_matches_oracle(margin, s, m):
    from fedfr_b200 import selfcheck as SC
    from oracle import partial_fc_oracle as O
    g = torch.Generator().manual_seed(3)
    B, C, E = 96, 1500, 128
    w = torch.randn(C, E, generator=g) * 0.01
    y = torch.randint(0, C, (B,), generator=g)
    y[5] = y[4]
    x = torch.nn.functional.normalize(torch.randn(B, E, generator=g) + 2.0 * torch.nn.functional.normalize(w[y]) * (torch.arange(B) % 3 == 0)[:, None])
    kind = 0 if margin == "cosface" else 1
    ref = O.forward_backward([x], [y], [w], C, s, m, margin=margin)
    got = SC.reference_step(x, y, w, s, m, kind, emulate=False, chunk=400)
    assert abs(float(got["loss"]) - float(ref.loss)) < 1e-5 * float(ref.loss)
    assert rel(got["dx_total"], ref.x_grad[0]) < 1e-5 and rel(got["dw"], ref.dw[0]) < 1e-5
    refb = O.forward_backward_bf16([x], [y], [w], C, s, m, margin=margin)
    gotb = SC.reference_step(x, y, w, s, m, kind, emulate=True, chunk=400)
    assert abs(float(gotb["loss"]) - float(refb.loss)) < 1e-5 * float(refb.loss)
    # same roundings, fp32 vs fp64 GEMMs in between: a few bf16 roundings flip to the other neighbour (one ulp = 2^-8 on a
    # row dominated by that element), nothing systematic (rms over the rows far below)
    for got_t, ref_t in ((gotb["dx_total"], refb.x_grad[0]), (gotb["dw"], refb.dw[0])):
        assert SC._rows_err(got_t, ref_t) < 1e-2 and SC._rows_err(got_t, ref_t, rms=True) < 1e-3
    # and the emulation stays within the bf16 tolerance of the exact arithmetic
    assert rel(refb.x_grad[0], ref.x_grad[0]) < 1e-2 and rel(refb.dw[0], ref.dw[0]) < 1e-2


def test_bf16_oracle_two_ranks_and_sampling():
    """forward_backward_bf16 follows forward_backward through the multi-rank / sampled bookkeeping (same indices, same
    label remap) and stays within 1e-2 of it."""
    from oracle import partial_fc_oracle as O
    g = torch.Generator().manual_seed(11)
    B, C, E, W = 32, 1001, 64, 2
    xs = [torch.nn.functional.normalize(torch.randn(B, E, generator=g)) for _ in range(W)]
    ys = [torch.randint(0, C, (B,), generator=g) for _ in range(W)]
    ws = [torch.randn(O.shard_geometry(C, W, r)[0], E, generator=g) * 0.01 for r in range(W)]
    perms = [np.random.default_rng(r).random(ws[r].shape[0], dtype=np.float32) for r in range(W)]
    a = O.forward_backward(xs, ys, ws, C, 64.0, 0.4, sample_rate=0.3, perms=perms)
    b = O.forward_backward_bf16(xs, ys, ws, C, 64.0, 0.4, sample_rate=0.3, perms=perms)
    assert abs(float(a.loss) - float(b.loss)) < 1e-3 * float(a.loss)
    for r in range(W):
        assert np.array_equal(a.index[r], b.index[r]) and np.array_equal(a.total_label[r], b.total_label[r])
        assert rel(b.x_grad[r], a.x_grad[r]) < 1e-2 and rel(b.dw[r], a.dw[r]) < 1e-2


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    try:
        import fedfr_b200
        from fedfr_b200 import selfcheck as SC
        from oracle_ops import OracleOps
        B, C, E = 24, 501, 64
        g = torch.Generator().manual_seed(100 + rank)
        head = fedfr_b200.PartialFC(rank, rank, world, B, False, fedfr_b200.CosFace(64.0, 0.4), C, sample_rate=0.5, embedding_size=E, prefix="/tmp",
                                    _ops=OracleOps())
        head.weight.copy_(torch.randn(head.num_local, E, generator=g) * 0.01)
        x = torch.nn.functional.normalize(torch.randn(B, E, generator=g))
        y = torch.randint(0, C, (B,), generator=g)
        opt = torch.optim.SGD([{"params": head.parameters()}], lr=0.1, momentum=0.9)
        xg, loss = head.forward_backward(y, x, opt)
        good = SC.check_head_step(head, y, x, xg, loss, dw_stride=7)
        bad_xg = xg.clone()
        if rank == 1:
            bad_xg[3] *= 1.5                                     # one wrong row on one rank must fail the job on every rank
        bad = SC.check_head_step(head, y, x, bad_xg, loss, dw_stride=7)
        ret[rank] = (good["ok"], good["ok_all_ranks"], good["fp32"]["dx_rel"], good["fp32"]["dw_rel"], bad["ok"], bad["ok_all_ranks"])
    finally:
        dist.barrier()
        dist.destroy_process_group()


def test_check_head_step_two_ranks_gloo():
    import __graft_entry__ as g
    g.build()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29741, ret), nprocs=2, join=True)
    for r in (0, 1):
        ok, ok_all, dx, dw, bad_ok, bad_all = ret[r]
        assert ok and ok_all and dx < 1e-5 and dw < 1e-5, dict(ret)
        assert not bad_all, dict(ret)
    assert ret[0][4] and not ret[1][4]          # rank 0's own numbers were fine; the all-rank verdict still fails
