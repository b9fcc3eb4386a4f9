"""SpreadOut_Module (server.py:48-63) on the fused kernels.

Pin: tests/golden/spreadout.npz holds the loss and FC.grad the UNMODIFIED reference class produced on CPU
(tests/golden/make_golden.py spreadout).  The oracle restatement is checked against it on CPU; the CUDA module is checked
against the golden and, at FedFR-like sizes, against the oracle under torch autograd in fp64."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
GOLD = np.load(os.path.join(HERE, "golden", "spreadout.npz"))


def _centres(n, e, seed):
    """Clustered class centres: groups of near-duplicates so that many pairs exceed the margin."""
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(max(n // 16, 1), e, generator=g)
    return (base[torch.arange(n) % len(base)] + 0.35 * torch.randn(n, e, generator=g)) * 0.05


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.parametrize("mode", ["sum", "mean"])
def test_oracle_matches_reference_golden(mode):
    from oracle import partial_fc_oracle as O
    fc = torch.from_numpy(GOLD["fc"]).clone().requires_grad_(True)
    loss = O.spreadout_loss(fc, float(GOLD[f"margin_{mode}"]), mode)
    loss.backward()
    assert abs(loss.item() - float(GOLD[f"loss_{mode}"])) <= 1e-5 * float(GOLD[f"loss_{mode}"])
    assert rel(fc.grad, GOLD[f"grad_{mode}"]) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["sum", "mean"])
def test_cuda_matches_reference_golden(mode):
    import __graft_entry__ as g
    g.build()
    import fedfr_b200
    mod = fedfr_b200.SpreadOut_Module(torch.from_numpy(GOLD["fc"]).cuda(), margin=float(GOLD[f"margin_{mode}"]), mode=mode)
    loss = mod()
    loss.backward()
    assert abs(loss.item() - float(GOLD[f"loss_{mode}"])) <= 1e-2 * float(GOLD[f"loss_{mode}"])
    assert rel(mod.FC.grad, GOLD[f"grad_{mode}"]) < 2e-2          # bf16 cosines right at the hinge: held to 2e-2


@pytest.mark.gpu
@pytest.mark.parametrize("n,e,margin,mode", [(6000, 512, 0.4, "sum"), (1000, 512, 0.7, "mean"), (777, 256, 0.4, "sum"), (300, 64, 0.2, "sum")])
def test_spreadout_vs_autograd(n, e, margin, mode):
    """FedFR scale: 6000 public + client centres, E = 512 (server.py:340-347); ragged sizes for the tile edges."""
    import __graft_entry__ as g
    g.build()
    import fedfr_b200
    from oracle import partial_fc_oracle as O
    fc = _centres(n, e, n + e)
    ref_fc = fc.double().requires_grad_(True)
    ref = O.spreadout_loss(ref_fc, margin, mode)
    ref.backward()
    assert ref.item() > 0
    mod = fedfr_b200.SpreadOut_Module(fc.clone().cuda(), margin=margin, mode=mode)
    loss = mod()
    (3.0 * loss).backward()                                        # grad_output != 1
    assert abs(loss.item() - ref.item()) <= 1e-2 * ref.item()
    assert rel(mod.FC.grad, 3.0 * ref_fc.grad) < 2e-2
    # the SGD loop of server.py:352-359 runs on the module unchanged
    opt = torch.optim.SGD(mod.parameters(), lr=0.1, momentum=0.9, weight_decay=5e-4)
    l0 = mod().item()
    for _ in range(3):
        opt.zero_grad()
        mod().backward()
        opt.step()
    assert mod().item() < l0
