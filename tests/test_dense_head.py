"""Autograd adapter for FedFR's dense head (client.py:69-74 + losses.py:23-29 + F.cross_entropy, client.py:430-435).

The checker is the reference's own formulation evaluated by torch autograd in fp32/fp64.  On CPU the adapter runs with
the oracle-backed ``ops`` provider (host logic: normalize forward/backward glue, grad_output scaling); on a B200 it
runs through the C ABI (``-m gpu``)."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def reference_loss(x, fc, y, s, m, kind):
    """client.py:69-74 -> losses.py:23-29 / 38-45 -> F.cross_entropy, written as in the reference."""
    cosine = torch.matmul(F.normalize(x), F.normalize(fc).t())
    onehot = F.one_hot(y, fc.shape[0]).to(cosine.dtype)
    if kind == "cosface":
        cosine = cosine - m * onehot
    else:
        theta = torch.acos(cosine.clamp(-1, 1))
        cosine = torch.where(onehot.bool(), torch.cos(theta + m), cosine)
    return F.cross_entropy(cosine * s, y)


def _case(B, C, E, seed, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, E, generator=g, dtype=dtype) * 3.0          # un-normalised backbone output
    fc = torch.randn(C, E, generator=g, dtype=dtype) * 0.01
    y = torch.randint(0, C, (B,), generator=g)
    x[::3] += 4.0 * fc[y[::3]] / fc[y[::3]].norm(dim=1, keepdim=True)   # some well-classified rows
    return x, fc, y


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-12))


@pytest.mark.parametrize("kind,s,m", [("cosface", 30.0, 0.4), ("arcface", 64.0, 0.5)])
def test_dense_head_host_logic_cpu(kind, s, m):
    import fedfr_b200
    from oracle_ops import OracleOps
    x, fc, y = _case(48, 211, 64, 3)
    margin = fedfr_b200.CosFace(s, m) if kind == "cosface" else fedfr_b200.ArcFace(s, m)
    xr, fr = x.double().requires_grad_(True), fc.double().requires_grad_(True)
    (2.5 * reference_loss(xr, fr, y, s, m, kind)).backward()
    xa, fa = x.clone().requires_grad_(True), fc.clone().requires_grad_(True)
    loss = fedfr_b200.margin_cross_entropy(xa, fa, y, margin, _ops=OracleOps())
    (2.5 * loss).backward()                                            # grad_output != 1 reaches both gradients
    assert abs(loss.item() - reference_loss(x.double(), fc.double(), y, s, m, kind).item()) < 1e-4 * loss.item()
    assert rel(xa.grad, xr.grad) < 1e-4 and rel(fa.grad, fr.grad) < 1e-4


@pytest.mark.parametrize("kind,s,m", [("cosface", 30.0, 0.4), ("arcface", 64.0, 0.5)])
def test_dense_head_vs_reference_golden_cpu(kind, s, m):
    """tests/golden/dense_head.npz: loss / x.grad / fc.grad of the UNMODIFIED reference classes (client.FC_module +
    losses.CosFace / ArcFace + F.cross_entropy + backward, tests/golden/make_golden.py dense).  The adapter's host logic
    (normalize forward / backward glue around the kernel provider) must reproduce them with the oracle-backed provider."""
    import numpy as np
    import fedfr_b200
    from oracle_ops import OracleOps
    gold = np.load(os.path.join(HERE, "golden", "dense_head.npz"))
    x = torch.from_numpy(gold["x"]).clone().requires_grad_(True)
    fc = torch.from_numpy(gold["fc"]).clone().requires_grad_(True)
    y = torch.from_numpy(gold["y"])
    margin = fedfr_b200.CosFace(s, m) if kind == "cosface" else fedfr_b200.ArcFace(s, m)
    loss = fedfr_b200.margin_cross_entropy(x, fc, y, margin, _ops=OracleOps())
    loss.backward()
    assert abs(loss.item() - float(gold[f"loss_{kind}"])) <= 1e-4 * float(gold[f"loss_{kind}"])
    assert rel(x.grad, torch.from_numpy(gold[f"dx_{kind}"])) < 1e-4
    assert rel(fc.grad, torch.from_numpy(gold[f"dfc_{kind}"])) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("check_mode,tol", [(True, 1e-4), (False, 1e-2)])
@pytest.mark.parametrize("kind,s,m", [("cosface", 30.0, 0.4), ("arcface", 64.0, 0.5)])
def test_dense_head_vs_torch_autograd(kind, s, m, check_mode, tol):
    """FedFR's scale: batch 256, 6100 classes (4000 local + public ids), E = 512."""
    import __graft_entry__ as g
    g.build()
    import fedfr_b200
    dev = torch.device("cuda:0")
    x, fc, y = _case(256, 6100, 512, 11)
    margin = fedfr_b200.CosFace(s, m) if kind == "cosface" else fedfr_b200.ArcFace(s, m)
    xr, fr = x.double().requires_grad_(True), fc.double().requires_grad_(True)
    ref = reference_loss(xr, fr, y, s, m, kind)
    ref.backward()
    head = fedfr_b200.MarginSoftmaxHead(6100, margin, 512, check_mode=check_mode).to(dev)
    head.fc.data.copy_(fc.to(dev))
    xa = x.to(dev).requires_grad_(True)
    loss = head(xa, y.to(dev))
    loss.backward()
    assert abs(loss.item() - ref.item()) <= tol * ref.item()
    assert rel(xa.grad.cpu(), xr.grad) < tol
    assert rel(head.fc.grad.cpu(), fr.grad) < tol
    # second forward/backward accumulates into .grad like any autograd op
    head(xa, y.to(dev)).backward()
    assert rel(head.fc.grad.cpu(), 2 * fr.grad) < tol


@pytest.mark.gpu
def test_two_heads_interleaved(kind="cosface", s=30.0, m=0.4):
    """Two heads of the same shape on one device share the library's step buffers: forward A, forward B, then one
    backward over both must still give each head its own gradients (A's backward rebuilds its operands)."""
    import __graft_entry__ as g
    g.build()
    import fedfr_b200
    dev = torch.device("cuda:0")
    xa, fca, ya = _case(256, 6100, 512, 21)
    xb, fcb, yb = _case(256, 6100, 512, 22)
    refs = []
    for x, fc, y in ((xa, fca, ya), (xb, fcb, yb)):
        xr, fr = x.double().requires_grad_(True), fc.double().requires_grad_(True)
        reference_loss(xr, fr, y, s, m, kind).backward()
        refs.append((xr.grad, fr.grad))
    heads, xs = [], []
    for x, fc in ((xa, fca), (xb, fcb)):
        h = fedfr_b200.MarginSoftmaxHead(6100, fedfr_b200.CosFace(s, m), 512).to(dev)
        h.fc.data.copy_(fc.to(dev))
        heads.append(h)
        xs.append(x.to(dev).requires_grad_(True))
    total = heads[0](xs[0], ya.to(dev)) + heads[1](xs[1], yb.to(dev))
    total.backward()
    for h, x, (gx, gw) in zip(heads, xs, refs):
        assert rel(x.grad.cpu(), gx) < 1e-2
        assert rel(h.fc.grad.cpu(), gw) < 1e-2
