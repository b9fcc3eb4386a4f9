"""World-size-2 run of fedfr_b200.PartialFC's HOST logic on gloo/CPU (kernels supplied by the oracle through
tests/oracle_ops.py) against the golden vectors the unmodified reference produced on 2 gloo ranks."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, name, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    try:
        import fedfr_b200
        from golden_util import Case, margin_of
        from oracle_ops import OracleOps
        case = Case(name)
        cfg = case.cfg
        head = fedfr_b200.PartialFC(rank, rank, world, cfg["batch"], False, margin_of(fedfr_b200, cfg), cfg["num_classes"],
                                    sample_rate=cfg["sample_rate"], embedding_size=cfg["emb"], prefix="/tmp", _ops=OracleOps())
        head.weight.copy_(case.weights[rank])
        head.weight_mom.zero_()
        opt = torch.optim.SGD([{"params": head.parameters()}], lr=cfg["lr"], momentum=0.9, weight_decay=5e-4)
        torch.manual_seed(cfg["seed"] * 1000 + rank)       # same RNG position as the reference run -> same perm
        errs = []
        for step in range(cfg["steps"]):
            x_grad, loss = head.forward_backward(case.labels[rank], case.features[rank], opt)
            errs.append(abs(float(loss) - float(case.get(rank, step, "loss"))))
            errs.append(float(np.abs(x_grad.numpy() - case.get(rank, step, "x_grad")).max()))
            errs.append(float(np.abs(head.sub_weight.grad.numpy() - case.get(rank, step, "dw")).max()))
            if case.has(rank, step, "index"):
                assert np.array_equal(head.index.numpy(), case.get(rank, step, "index"))
            opt.step()
            head.update()
            opt.zero_grad()
            errs.append(float(np.abs(head.weight.numpy() - case.get(rank, step, "weight_after")).max()))
            errs.append(float(np.abs(head.weight_mom.numpy() - case.get(rank, step, "mom_after")).max()))
        ret[rank] = max(errs)
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("name,port", [("w2_sr1_ragged", 29731), ("w2_sr03", 29732), ("w2_arc_sr03", 29733)])
def test_two_rank_host_logic_matches_reference(name, port):
    import __graft_entry__ as g
    g.build()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, name, port, ret), nprocs=2, join=True)
    assert set(ret.keys()) == {0, 1}
    assert max(ret.values()) < 5e-5, dict(ret)


@pytest.mark.parametrize("name", ["w1_sr1_small", "w1_sr01", "w1_sr_pos_overflow", "w1_arc_small"])
def test_single_rank_host_logic_matches_reference(name):
    """W=1 needs no process group (the reference does; the drop-in accepts both)."""
    import __graft_entry__ as g
    g.build()
    sys.path.insert(0, HERE)
    import fedfr_b200
    from golden_util import Case, margin_of
    from oracle_ops import OracleOps
    case = Case(name)
    cfg = case.cfg
    head = fedfr_b200.PartialFC(0, 0, 1, cfg["batch"], False, margin_of(fedfr_b200, cfg), cfg["num_classes"],
                                sample_rate=cfg["sample_rate"], embedding_size=cfg["emb"], prefix="/tmp", _ops=OracleOps())
    head.weight.copy_(case.weights[0])
    opt = torch.optim.SGD([{"params": head.parameters()}], lr=cfg["lr"], momentum=0.9, weight_decay=5e-4)
    assert [p is head.sub_weight for p in head.parameters()] == [True]
    assert list(head.state_dict().keys()) == ["sub_weight"]
    torch.manual_seed(cfg["seed"] * 1000)
    # ArcFace: the reference runs acos_/cos_ over every logit (an identity up to ~1e-7 per element), see test_oracle_golden
    atol = 1e-5 if cfg.get("loss", "cosface") == "arcface" else 2e-6
    for step in range(cfg["steps"]):
        x_grad, loss = head.forward_backward(case.labels[0], case.features[0], opt)
        assert abs(float(loss) - float(case.get(0, step, "loss"))) < 5e-5
        np.testing.assert_allclose(x_grad.numpy(), case.get(0, step, "x_grad"), rtol=2e-4, atol=atol)
        np.testing.assert_allclose(head.sub_weight.grad.numpy(), case.get(0, step, "dw"), rtol=2e-4, atol=atol)
        if case.has(0, step, "index"):
            np.testing.assert_array_equal(head.index.numpy(), case.get(0, step, "index"))
        assert opt.param_groups[-1]["params"][0] is head.sub_weight
        assert opt.state[head.sub_weight]["momentum_buffer"] is head.sub_weight_mom
        opt.step()
        head.update()
        opt.zero_grad()
        np.testing.assert_allclose(head.weight.numpy(), case.get(0, step, "weight_after"), rtol=3e-4, atol=3e-6)
        np.testing.assert_allclose(head.weight_mom.numpy(), case.get(0, step, "mom_after"), rtol=3e-4, atol=3e-6)


@pytest.mark.parametrize("name", ["w1_sr1_small", "w1_sr01"])
def test_fused_step_host_logic_matches_reference(name):
    """PartialFC.step(optimizer) == optimizer.step() + update() of the reference run (weights and momentum after each step)."""
    import __graft_entry__ as g
    g.build()
    sys.path.insert(0, HERE)
    import fedfr_b200
    from golden_util import Case, margin_of
    from oracle_ops import OracleOps
    case = Case(name)
    cfg = case.cfg
    head = fedfr_b200.PartialFC(0, 0, 1, cfg["batch"], False, margin_of(fedfr_b200, cfg), cfg["num_classes"],
                                sample_rate=cfg["sample_rate"], embedding_size=cfg["emb"], prefix="/tmp", _ops=OracleOps())
    head.weight.copy_(case.weights[0])
    opt = torch.optim.SGD([{"params": head.parameters()}], lr=cfg["lr"], momentum=0.9, weight_decay=5e-4)
    torch.manual_seed(cfg["seed"] * 1000)
    for step in range(cfg["steps"]):
        head.forward_backward(case.labels[0], case.features[0], opt)
        head.step(opt)                          # instead of opt.step(); head.update()
        opt.zero_grad()
        np.testing.assert_allclose(head.weight.numpy(), case.get(0, step, "weight_after"), rtol=3e-4, atol=3e-6)
        np.testing.assert_allclose(head.weight_mom.numpy(), case.get(0, step, "mom_after"), rtol=3e-4, atol=3e-6)
