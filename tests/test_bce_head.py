"""Cosine head of the personalised branch (client.py:25-60 BCE_module + losses.py:4-15 BCE_loss + loss.backward()).

Golden: tests/golden/bce_head.npz -- logits, gt, loss and every gradient of the UNMODIFIED reference classes on CPU
(tests/golden/make_golden.py bce).  CPU: the oracle against the golden, and fedfr_b200.BCE_module's host logic (converter,
autograd glue, state_dict) with the oracle-backed ops.  GPU: the two fused launches through the C ABI against the golden
and against torch autograd of the reference formulation, fp32 tolerance 1e-4 relative (BASELINE north_star check mode)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
TOL = 1e-4


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__ as g
    g.build()
    import fedfr_b200
    return fedfr_b200


@pytest.fixture(scope="module")
def golden():
    z = np.load(os.path.join(HERE, "golden", "bce_head.npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-12))


def bce_loss(logits, gts, r=30.0, lam=0.7):
    """losses.py:4-15 written without the in-place edits."""
    pos = (lam / r) * torch.log(1 + torch.exp(-logits) + 1e-8)
    neg = ((1 - lam) / r) * torch.log(1 + torch.exp(logits) + 1e-8)
    return torch.where(gts, pos, neg).sum(dim=1).mean()


def reference_head(feat, weight, bias, labels, m, r, t):
    """client.py:47-57 written with torch ops (the checker for shapes the golden does not cover)."""
    n_class = weight.shape[0]
    cosine = torch.matmul(F.normalize(feat), F.normalize(weight).t())
    lab = labels.clone()
    lab[lab >= n_class] = n_class
    gt = F.one_hot(lab, n_class + 1)[:, :-1].bool()
    g = 2 * ((cosine + 1) / 2).pow(t) - 1
    return torch.where(gt, r * (g - m), r * (g + m)) + bias[None, :], gt


def _load_module(pkg, golden, device="cpu", ops=None):
    C, E = golden["sd/weight"].shape
    mod = pkg.BCE_module(E, C, 1, _ops=ops)
    missing = mod.load_state_dict({k[3:]: v for k, v in golden.items() if k.startswith("sd/")})   # bce_module.pth layout
    assert not missing.missing_keys and not missing.unexpected_keys
    return mod.to(device)


def _check_against_golden(mod, golden, device, tol):
    x = golden["x"].detach().clone().to(device).requires_grad_(True)
    logits, gts = mod(x, golden["y"].to(device))
    assert gts.dtype == torch.bool and torch.equal(gts.cpu(), golden["gt"])
    assert rel(logits.detach(), golden["logits"]) < tol
    loss = 10 * bce_loss(logits, gts)
    loss.backward()
    assert abs(float(loss.detach()) - float(golden["loss"])) < tol * abs(float(golden["loss"]))
    assert rel(x.grad, golden["dx"]) < tol
    for k, p in mod.named_parameters():
        assert rel(p.grad, golden["grad/" + k]) < tol, k


# ----------------------------------------------------------------------------------------------- CPU

def test_oracle_matches_reference_golden(golden):
    from oracle import bce_head_oracle as O
    sd = {k[3:]: v.double() for k, v in golden.items() if k.startswith("sd/")}
    feat = golden["x"].double() @ sd["converter.0.weight"].t() + sd["converter.0.bias"]
    logits, gt, cosine, nf, nw = O.forward(feat, sd["weight"], sd["bias"], golden["y"], 0.4, 30.0, 3)
    assert torch.equal(gt, golden["gt"]) and rel(logits, golden["logits"]) < 1e-6
    lg = logits.clone().requires_grad_(True)
    (10 * bce_loss(lg, gt)).backward()
    dfeat, dweight, dbias = O.backward(feat, sd["weight"], cosine, nf, nw, lg.grad, 30.0, 3)
    assert rel(dweight, golden["grad/weight"]) < 1e-5 and rel(dbias, golden["grad/bias"]) < 1e-5
    assert rel(dfeat @ sd["converter.0.weight"], golden["dx"]) < 1e-5
    assert rel(dfeat.t() @ golden["x"].double(), golden["grad/converter.0.weight"]) < 1e-5


def test_target_column_rule():
    from oracle import bce_head_oracle as O
    labels = torch.tensor([0, 4, 5, 9, -1, -2, -6, -7])
    want = F.one_hot(torch.tensor([0, 4, 5, 5, 5, 4, 0, 5]), 6)[:, :-1].bool()    # python indexing into [B, 6], last col dropped
    want[7] = False                                                               # -7 would raise in torch: no column
    col = O.target_col(labels, 5)
    gt = torch.zeros(8, 5, dtype=torch.bool)
    rows = torch.nonzero(col >= 0, as_tuple=True)[0]
    gt[rows, col[rows]] = True
    assert torch.equal(gt, want)


def test_module_host_logic_cpu(pkg, golden):
    from oracle.bce_head_oracle import OracleBceOps
    mod = _load_module(pkg, golden, ops=OracleBceOps())
    assert list(mod.state_dict().keys()) == [k[3:] for k in golden if k.startswith("sd/")]      # the reference's key order
    assert sorted(mod.state_dict().keys()) == ["bias", "converter.0.bias", "converter.0.weight", "weight"]
    _check_against_golden(mod, golden, "cpu", 1e-5)


def test_module_surface(pkg):
    from oracle.bce_head_oracle import OracleBceOps
    mod = pkg.BCE_module(16, 5, _ops=OracleBceOps())
    assert torch.equal(mod.converter[0].weight.data, torch.eye(16)) and float(mod.converter[0].bias.abs().max()) == 0
    assert tuple(mod.weight.shape) == (5, 16) and float(mod.bias.abs().max()) == 0 and (mod.m, mod.r, mod.n_class, mod.hidden) == (0.4, 30.0, 5, 16)
    fc = torch.randn(5, 16)
    mod.initialize(fc)                                              # client.py:59-60
    assert torch.equal(mod.weight.data, fc) and mod.weight.data.data_ptr() != fc.data_ptr()
    x = torch.randn(3, 16)
    logits, gt = mod(x.detach(), torch.tensor([1, 7, 4]))           # detach=True branch of Branch_model (client.py:94-95)
    assert gt.tolist() == [[False, True, False, False, False], [False] * 5, [False, False, False, False, True]]
    logits.sum().backward()
    assert mod.weight.grad is not None and mod.converter[0].weight.grad is not None
    with pytest.raises(NotImplementedError):
        pkg.BCE_module(16, 5, converter_layer=2)
    with pytest.raises(RuntimeError):
        pkg.BCE_module(16, 5)(x, torch.tensor([1, 2, 3]))           # CPU tensors without the test provider: no fallback


def test_reference_bce_loss_runs_on_our_logits(pkg, golden):
    """losses.BCE_loss edits the logits in place (losses.py:10-11); the head's output must allow that."""
    from oracle.bce_head_oracle import OracleBceOps
    mod = _load_module(pkg, golden, ops=OracleBceOps())
    x = golden["x"].detach().clone().requires_grad_(True)
    logits, gts = mod(x, golden["y"])
    logits[gts] = (0.7 / 30) * torch.log(1 + torch.exp(-1 * logits[gts]) + 1e-8)
    logits[~gts] = (0.3 / 30) * torch.log(1 + torch.exp(1 * logits[~gts]) + 1e-8)
    loss = 10 * torch.mean(torch.sum(logits, dim=1))
    loss.backward()
    assert abs(float(loss.detach()) - float(golden["loss"])) < 1e-5 and rel(x.grad, golden["dx"]) < 1e-5


# ----------------------------------------------------------------------------------------------- GPU

@pytest.mark.gpu
def test_gpu_matches_reference_golden(pkg, golden):
    _check_against_golden(_load_module(pkg, golden, "cuda:0"), golden, "cuda:0", TOL)


@pytest.mark.gpu
@pytest.mark.parametrize("B,C,E,t,detach", [(256, 100, 512, 3, False), (64, 300, 512, 3, True), (7, 3, 33, 2, False), (5, 1, 1024, 4, False)])
def test_gpu_matches_reference_formulation(pkg, B, C, E, t, detach):
    g = torch.Generator().manual_seed(B + C)
    feat = torch.randn(B, E, generator=g) * 2
    weight = torch.randn(C, E, generator=g) * 0.01
    bias = torch.randn(C, generator=g) * 0.1
    y = torch.randint(0, C + C // 3 + 1, (B,), generator=g)
    fr, wr, br = (v.double().requires_grad_(True) for v in (feat, weight, bias))
    lr, gr = reference_head(fr, wr, br, y, 0.4, 30.0, t)
    bce_loss(lr, gr).backward()

    mod = pkg.BCE_module(E, C, 1, t=t).to("cuda:0")
    with torch.no_grad():
        mod.weight.copy_(weight)
        mod.bias.copy_(bias)
    x = feat.to("cuda:0").requires_grad_(not detach)
    logits, gts = mod(x, y.to("cuda:0"))                            # identity converter: feat == x
    assert torch.equal(gts.cpu(), gr) and rel(logits.detach(), lr.detach()) < TOL
    bce_loss(logits, gts).backward()
    assert rel(mod.weight.grad, wr.grad) < TOL and rel(mod.bias.grad, br.grad) < TOL
    if detach:
        assert x.grad is None
    else:
        assert rel(x.grad, fr.grad) < TOL
        assert rel(mod.converter[0].weight.grad, fr.grad.t() @ feat.double()) < TOL


@pytest.mark.gpu
def test_gpu_empty_batch(pkg):
    mod = pkg.BCE_module(64, 5).to("cuda:0")
    logits, gts = mod(torch.zeros(0, 64, device="cuda:0"), torch.zeros(0, dtype=torch.long, device="cuda:0"))
    assert tuple(logits.shape) == (0, 5) and tuple(gts.shape) == (0, 5)
    logits.sum().backward()
    assert float(mod.weight.grad.abs().max()) == 0.0 and float(mod.bias.grad.abs().max()) == 0.0
