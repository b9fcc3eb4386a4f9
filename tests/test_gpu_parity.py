"""Parity of the CUDA path (through the C ABI, via fedfr_b200.PartialFC / FedPavg) with the reference.

* golden vectors produced by the unmodified reference (tests/golden/*.npz)           -> both kernel paths
* the CPU oracle on seeded inputs at sizes it finishes in seconds (incl. BASELINE c1)  -> both kernel paths
* BASELINE.json full sizes (1M classes; 250k-class sampled shard; 40-client FedAvg)    -> size-independent
  properties + a chunked device-side restatement

Tolerances (BASELINE.json north_star): integer / index work bit-exact; loss, x_grad, dw within 1e-2 relative
on the bf16 tensor path and 1e-4 in fp32 check mode.  Run with ``pytest -m gpu`` on a B200.
"""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

TOL = {False: 1e-2, True: 1e-4}      # check_mode -> relative tolerance


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__ as g
    g.build()
    import fedfr_b200
    return fedfr_b200


from golden_util import margin_of  # noqa: E402  (tests/ is on sys.path via conftest)


def rel(a, b):
    """Relative L2 error with an absolute floor (1e-7 rms) so vanishing gradients do not divide by ~0."""
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    return float((a - b).norm() / (b.norm() + 1e-7 * b.numel() ** 0.5))


def _make_head(pkg, cfg, weight, check_mode, rank=0, world=1):
    head = pkg.PartialFC(rank, 0, world, cfg["batch"], False, margin_of(pkg, cfg), cfg["num_classes"],
                         sample_rate=cfg["sample_rate"], embedding_size=cfg["emb"], prefix="/tmp", check_mode=check_mode)
    head.weight.copy_(weight.to(head.device))
    head.weight_mom.zero_()
    return head


@pytest.mark.parametrize("check_mode", [True, False])
@pytest.mark.parametrize("name", ["w1_sr1_small", "w1_sr1_s30", "w1_sr01", "w1_sr_pos_overflow", "w1_arc_small"])
def test_golden_single_rank(pkg, name, check_mode, monkeypatch):
    from golden_util import Case, margin_of
    case = Case(name)
    cfg = case.cfg
    if not check_mode and cfg["emb"] not in (64, 128, 256, 512):
        pytest.skip("tensor path supports E in {64,128,256,512}")
    head = _make_head(pkg, cfg, case.weights[0], check_mode)
    dev = head.device
    opt = torch.optim.SGD([{"params": head.parameters()}], lr=cfg["lr"], momentum=0.9, weight_decay=5e-4)
    tol = TOL[check_mode]
    real_rand = torch.rand
    for step in range(cfg["steps"]):
        if case.has(0, step, "perm"):       # feed the reference's own torch.rand draw (CPU generator) to the index kernel
            perm = torch.from_numpy(case.get(0, step, "perm")).to(dev)
            monkeypatch.setattr(torch, "rand", lambda *a, **k: perm.clone())
        x_grad, loss = head.forward_backward(case.labels[0].to(dev), case.features[0].to(dev), opt)
        monkeypatch.setattr(torch, "rand", real_rand)
        if case.has(0, step, "index"):
            assert np.array_equal(head.index.cpu().numpy(), case.get(0, step, "index")), "sampled index must be bit-exact"
        assert abs(float(loss) - float(case.get(0, step, "loss"))) <= tol * max(1.0, abs(float(case.get(0, step, "loss"))))
        assert rel(x_grad, case.get(0, step, "x_grad")) < tol
        assert rel(head.sub_weight.grad, case.get(0, step, "dw")) < tol
        opt.step()
        head.update()
        opt.zero_grad()
        # compare the update (lr * momentum buffer), not the weights it is added to
        w_ref, w0 = case.get(0, step, "weight_after"), (case.weights[0].numpy() if step == 0 else prev_w)
        assert rel(head.weight.cpu().numpy() - w0, w_ref - w0) < 2 * tol
        prev_w = w_ref
        head.weight.copy_(torch.from_numpy(w_ref).to(dev))                   # keep the trajectories aligned
        head.weight_mom.copy_(torch.from_numpy(case.get(0, step, "mom_after")).to(dev))
        if int(cfg["sample_rate"]) == 1:
            assert head.sub_weight.data_ptr() == head.weight.data_ptr()       # aliasing contract, partial_fc.py:64-67


@pytest.mark.parametrize("check_mode", [True, False])
def test_c1_config_vs_golden(pkg, check_mode):
    """BASELINE.json configs[0]: B=128, C=10k, E=512, sample_rate=1 (reference run on CPU gloo)."""
    from golden_util import Case, margin_of
    case = Case("c1_b128_c10k")
    head = _make_head(pkg, case.cfg, case.weights[0], check_mode)
    dev = head.device
    opt = torch.optim.SGD([{"params": head.parameters()}], lr=case.cfg["lr"], momentum=0.9, weight_decay=5e-4)
    x_grad, loss = head.forward_backward(case.labels[0].to(dev), case.features[0].to(dev), opt)
    tol = TOL[check_mode]
    assert abs(float(loss) - float(case.get(0, 0, "loss"))) <= tol * float(case.get(0, 0, "loss"))
    assert rel(x_grad, case.get(0, 0, "x_grad")) < tol
    assert rel(head.sub_weight.grad[::97], case.get(0, 0, "dw_rows")) < tol
    assert abs(float(head.sub_weight.grad.double().norm()) / float(case.get(0, 0, "dw_norm")) - 1) < tol


@pytest.mark.parametrize("check_mode", [True, False])
@pytest.mark.parametrize("B,C,E,s", [(64, 1000, 512, 64.0), (200, 4097, 256, 30.0), (512, 20000, 512, 64.0), (3, 5, 64, 64.0)])
def test_vs_oracle(pkg, B, C, E, s, check_mode):
    """Seeded inputs, the oracle as checker; ragged shapes (rows / classes not multiples of any tile)."""
    from oracle import partial_fc_oracle as O
    g = torch.Generator().manual_seed(B * 7 + C)
    x = torch.nn.functional.normalize(torch.randn(B, E, generator=g))
    y = torch.randint(0, C, (B,), generator=g)
    w = torch.randn(C, E, generator=g) * 0.01
    ref = O.forward_backward([x], [y], [w], C, s, 0.4)
    cfg = dict(batch=B, num_classes=C, emb=E, s=s, m=0.4, sample_rate=1.0)
    head = _make_head(pkg, cfg, w, check_mode)
    opt = torch.optim.SGD([{"params": head.parameters()}], lr=0.1, momentum=0.9)
    x_grad, loss = head.forward_backward(y.to(head.device), x.to(head.device), opt)
    tol = TOL[check_mode]
    assert abs(float(loss) - float(ref.loss)) <= tol * float(ref.loss)
    assert rel(x_grad, ref.x_grad[0]) < tol
    assert rel(head.sub_weight.grad, ref.dw[0]) < tol
    # second call without zero_grad accumulates into .grad (torch semantics of logits.backward, partial_fc.py:168)
    head.forward_backward(y.to(head.device), x.to(head.device), opt)
    assert rel(head.sub_weight.grad, 2 * ref.dw[0]) < tol


@pytest.mark.parametrize("check_mode", [True, False])
@pytest.mark.parametrize("B,C,E,s", [(200, 4097, 256, 30.0), (512, 20000, 512, 64.0)])
def test_arcface_vs_oracle(pkg, B, C, E, s, check_mode):
    """ArcFace margin (losses.py:32-45) through the same fused kernels; labels drawn from few classes so that trained-like
    (large) target cosines occur: the slope sin(theta + m) / sin(theta) is exercised away from theta = pi / 2."""
    from oracle import partial_fc_oracle as O
    g = torch.Generator().manual_seed(B * 11 + C)
    w = torch.randn(C, E, generator=g) * 0.01
    y = torch.randint(0, C, (B,), generator=g)
    x = torch.nn.functional.normalize(torch.randn(B, E, generator=g) + 3.0 * torch.nn.functional.normalize(w[y]) * (torch.arange(B) % 3 == 0)[:, None])
    ref = O.forward_backward([x], [y], [w], C, s, 0.5, margin="arcface")
    cfg = dict(batch=B, num_classes=C, emb=E, s=s, m=0.5, sample_rate=1.0, loss="arcface")
    head = _make_head(pkg, cfg, w, check_mode)
    opt = torch.optim.SGD([{"params": head.parameters()}], lr=0.1, momentum=0.9)
    x_grad, loss = head.forward_backward(y.to(head.device), x.to(head.device), opt)
    tol = TOL[check_mode]
    assert abs(float(loss) - float(ref.loss)) <= tol * float(ref.loss)
    assert rel(x_grad, ref.x_grad[0]) < tol
    assert rel(head.sub_weight.grad, ref.dw[0]) < tol


def test_hard_sample_clamp(pkg):
    """A row whose target probability underflows must hit the 1e-30 clamp (partial_fc.py:162)."""
    from oracle import partial_fc_oracle as O
    B, C, E = 8, 300, 64
    g = torch.Generator().manual_seed(5)
    w = torch.randn(C, E, generator=g)
    x = torch.nn.functional.normalize(torch.randn(B, E, generator=g))
    y = torch.randint(0, C, (B,), generator=g)
    x[0] = -torch.nn.functional.normalize(w[y[0]], dim=0)          # cos = -1 to its own class
    ref = O.forward_backward([x], [y], [w], C, 64.0, 0.4)
    head = _make_head(pkg, dict(batch=B, num_classes=C, emb=E, s=64.0, m=0.4, sample_rate=1.0), w, True)
    opt = torch.optim.SGD([{"params": head.parameters()}], lr=0.1)
    _, loss = head.forward_backward(y.to(head.device), x.to(head.device), opt)
    assert abs(float(loss) - float(ref.loss)) < 1e-3 * float(ref.loss)
    assert float(ref.loss) > 69.0 / B


def test_sampling_exact_vs_torch_cuda(pkg):
    """Index kernel == torch.unique/topk/sort/searchsorted on the device (partial_fc.py:94-104), incl. forced
    ties at the threshold, positives > num_sample, and BASELINE config 4's shard (250k classes, 25k sampled)."""
    from fedfr_b200.ops_cuda import CudaOps
    dev = torch.device("cuda:0")
    ops = CudaOps(dev)
    torch.manual_seed(3)
    for (nl, k, nlab, quant) in [(250000, 25000, 4096, 0), (250000, 25000, 4096, 4096), (1000, 100, 64, 16), (300, 30, 64, 0),
                                 (77, 0, 16, 0), (1 << 20, 104857, 512, 0), (50, 50, 10, 0), (64, 10, 8, 2)]:
        lab = torch.randint(-1, nl, (nlab,), device=dev)
        if nl == 300:
            lab = torch.randint(0, 60, (nlab,), device=dev) * 5
        perm = torch.rand(nl, device=dev)
        if quant:
            perm = torch.floor(perm * quant) / quant
        pos = torch.unique(lab[lab >= 0], sorted=True)
        p2 = perm.clone()
        if k - pos.numel() >= 0:
            p2[pos] = 2.0
            ref_idx = torch.topk(p2, k=k)[1].sort()[0]
        else:
            ref_idx = pos
        ref_lab = lab.clone()
        ref_lab[lab >= 0] = torch.searchsorted(ref_idx, lab[lab >= 0])
        mine = lab.clone()
        idx = ops.sample(mine, perm.clone(), k)
        assert idx.shape == ref_idx.shape and bool((idx == ref_idx).all()), (nl, k, nlab, quant)
        assert bool((mine == ref_lab).all())
        # invariants
        assert bool((idx[1:] > idx[:-1]).all()) and bool(torch.isin(pos, idx).all())
        assert idx.numel() == max(k, pos.numel())


def test_sampling_oracle_small(pkg):
    from fedfr_b200.ops_cuda import CudaOps
    from oracle import partial_fc_oracle as O
    ops = CudaOps(torch.device("cuda:0"))
    rng = np.random.default_rng(0)
    for _ in range(20):
        nl = int(rng.integers(20, 3000)); k = int(rng.integers(0, nl)); nlab = int(rng.integers(1, 200))
        y = rng.integers(-1, nl, size=nlab)
        perm = (rng.integers(0, 64, size=nl) / 64).astype(np.float32) if rng.random() < 0.5 else rng.random(nl, dtype=np.float32)
        ref_idx = O.sample_index(y, perm, k)
        ref_lab = O.relabel_to_sample(y, ref_idx)
        lab = torch.from_numpy(y).cuda()
        idx = ops.sample(lab, torch.from_numpy(perm).cuda(), k)
        assert np.array_equal(idx.cpu().numpy(), ref_idx) and np.array_equal(lab.cpu().numpy(), ref_lab)


def test_fedavg_golden_bit_exact(pkg):
    from golden_util import GOLDEN
    z = np.load(os.path.join(GOLDEN, "fedavg.npz"))
    K = int(z["K"])
    keys = sorted({k.split("/", 1)[1] for k in z.files if k.startswith("in0/")})
    models = [{k: torch.from_numpy(z[f"in{i}/{k}"]).cuda() for k in keys} for i in range(K)]
    w = [int(v) for v in z["weights"]]
    out = pkg.FedPavg(models, w)
    for k in keys:
        assert out[k].dtype == torch.float32
        assert np.array_equal(out[k].cpu().numpy(), z[f"out/{k}"]), k
    # CPU-resident state_dicts (how FedFR holds them, client.py:469) go through the same kernel
    out2 = pkg.FedPavg([{k: v.cpu() for k, v in m.items()} for m in models], w)
    for k in keys:
        assert np.array_equal(out2[k].cpu().numpy(), z[f"out/{k}"]), k
    fcs = [torch.from_numpy(z[f"fc_in{i}"]).cuda() for i in range(K)]
    old = torch.from_numpy(z["fc_old"]).cuda()
    assert np.array_equal(pkg.FedAvg_on_FC(old, fcs, w, 1).cpu().numpy(), z["fc_out_p1"])
    assert np.array_equal(pkg.FedAvg_on_FC(old, fcs, w, 0.7).cpu().numpy(), z["fc_out_p07"])
    # load_state_dict truncation of the fp32 counter back to int64 (server.py:333)
    bn = torch.nn.BatchNorm1d(8)
    cnt = out["bn.num_batches_tracked"]
    bn.num_batches_tracked.copy_(cnt)
    assert int(bn.num_batches_tracked) == int(float(cnt))


def test_fedavg_full_size_properties(pkg):
    """BASELINE config 5 shape: K=40 clients x (43.6M-element backbone-sized tensor + Linear(512,512) + int64 counters)."""
    dev = torch.device("cuda:0")
    K = 40
    n_big = 43_629_071 - 79
    torch.manual_seed(9)
    base = torch.randn(n_big, device=dev)
    models = []
    for i in range(K):
        models.append({"backbone.flat": base + 0.01 * torch.randn(n_big, device=dev), "converter.weight": torch.randn(512, 512, device=dev),
                       "converter.bias": torch.randn(512, device=dev), "bn.num_batches_tracked": torch.tensor(100 * i + 7, device=dev)})
    w = [6000 + 37 * i for i in range(K)]
    out = pkg.FedPavg(models, w)
    wn = [x / sum(w) for x in w]
    ref = 0
    for i in range(K):
        ref += wn[i] * models[i]["backbone.flat"]
    assert bool((out["backbone.flat"] == ref).all())                # bit exact at full size
    # idempotence: averaging K copies of the average returns it (weights sum to 1 up to rounding)
    again = pkg.FedPavg([out] * 4, [1, 1, 1, 1])
    assert rel(again["backbone.flat"], out["backbone.flat"]) < 1e-6
    # linearity in the weights: one-hot weights select a client exactly
    sel = pkg.FedPavg(models[:3], [0, 1, 0])
    assert bool((sel["backbone.flat"] == models[1]["backbone.flat"]).all())


def _chunked_device_reference(x, w, y, s, m, chunk=65536):
    """Device-side restatement for full sizes (fp32 torch ops on bf16-rounded operands, class chunks)."""
    xb = x.to(torch.bfloat16).float()
    n = w.norm(dim=1, keepdim=True).clamp_min(1e-12)
    wb = (w / n).to(torch.bfloat16).float()
    B, C = x.shape[0], w.shape[0]
    M = torch.full((B,), -float("inf"), device=x.device)
    for c0 in range(0, C, chunk):
        z = xb @ wb[c0:c0 + chunk].t()
        hit = (y >= c0) & (y < c0 + chunk)
        z[hit, y[hit] - c0] -= m
        M = torch.maximum(M, (z * s).max(dim=1)[0])
    S = torch.zeros(B, device=x.device, dtype=torch.float64)
    tz = torch.zeros(B, device=x.device)
    for c0 in range(0, C, chunk):
        z = xb @ wb[c0:c0 + chunk].t()
        hit = (y >= c0) & (y < c0 + chunk)
        z[hit, y[hit] - c0] -= m
        z *= s
        tz[hit] = z[hit, y[hit] - c0]
        S += torch.exp(z - M[:, None]).double().sum(dim=1)
    loss = -(torch.exp(tz - M).double() / S).clamp_min(1e-30).log().mean()
    dx = torch.zeros_like(x, dtype=torch.float64)
    return M, S, loss, xb, wb, n


def test_full_size_c3_properties(pkg):
    """BASELINE configs[2] at W=1: B=512, 1M classes.  Checks loss and x_grad against a chunked device
    restatement, dw on a strided subset of rows, and sum_j G_ij = 0 via  x_grad . x  identities."""
    dev = torch.device("cuda:0")
    B, C, E, s, m = 512, 1_000_000, 512, 64.0, 0.4
    torch.manual_seed(100)
    head = pkg.PartialFC(0, 0, 1, B, False, pkg.CosFace(s=s, m=m), C, embedding_size=E, prefix="/tmp")
    x = torch.nn.functional.normalize(torch.randn(B, E, device=dev))
    y = torch.randint(0, C, (B,), device=dev)
    opt = torch.optim.SGD([{"params": head.parameters()}], lr=0.1, momentum=0.9)
    x_grad, loss = head.forward_backward(y, x, opt)
    M, S, loss_ref, xb, wb, n = _chunked_device_reference(x, head.weight, y, s, m)
    assert abs(float(loss) - float(loss_ref)) < 1e-3 * float(loss_ref)
    # gradient pieces on the device in fp64 for a subset of classes (every 1009th row + the targets)
    rows = torch.unique(torch.cat([torch.arange(0, C, 1009, device=dev), y]))
    z = (xb @ wb[rows].t())
    hit = (rows[None, :] == y[:, None])
    z = (z - m * hit) * s
    G = (torch.exp(z - M[:, None]).double() / S[:, None] - hit.double()) * (s / B)
    dwh = G.t() @ xb.double()
    wr = wb[rows].double()
    dw_ref = (dwh - wr * (wr * dwh).sum(1, keepdim=True)) / n[rows].double()
    assert rel(head.sub_weight.grad[rows], dw_ref) < 1e-2
    # x_grad: full reference in chunks
    dx = torch.zeros(B, E, device=dev, dtype=torch.float64)
    for c0 in range(0, C, 65536):
        zc = xb @ wb[c0:c0 + 65536].t()
        h = (y >= c0) & (y < c0 + 65536)
        zc[h, y[h] - c0] -= m
        Gc = torch.exp(zc * s - M[:, None]).double() / S[:, None]
        Gc[h, y[h] - c0] -= 1.0
        dx += (Gc * (s / B)) @ wb[c0:c0 + 65536].double()
    assert rel(x_grad, dx) < 1e-2
    assert head.sub_weight.grad.shape == (C, E) and bool(torch.isfinite(head.sub_weight.grad).all())


@pytest.mark.parametrize("sample_rate", [1.0, 0.25])
@pytest.mark.parametrize("nesterov,dampening", [(False, 0.0), (True, 0.0), (False, 0.1)])
def test_fused_step_matches_torch_sgd(pkg, sample_rate, nesterov, dampening, monkeypatch):
    """PartialFC.step(opt) vs torch.optim.SGD.step() + update() on identical state: same weights and momentum (the
    kernel mirrors torch's foreach arithmetic), and the pre-normalised operands it leaves behind give the same next
    forward_backward as the ordinary path."""
    B, C, E = 64, 3000, 256
    g = torch.Generator().manual_seed(99)
    w = torch.randn(C, E, generator=g) * 0.01
    x = torch.nn.functional.normalize(torch.randn(B, E, generator=g))
    y = torch.randint(0, C, (B,), generator=g)
    cfg = dict(batch=B, num_classes=C, emb=E, s=64.0, m=0.4, sample_rate=sample_rate)
    heads = [_make_head(pkg, cfg, w, False) for _ in range(2)]
    dev = heads[0].device
    opts = [torch.optim.SGD([{"params": h.parameters()}], lr=0.05, momentum=0.9, weight_decay=5e-4, nesterov=nesterov, dampening=dampening)
            for h in heads]
    perm = torch.rand(C, generator=g).to(dev)
    real_rand = torch.rand
    outs = []
    for it in range(3):
        for k, (h, o) in enumerate(zip(heads, opts)):
            monkeypatch.setattr(torch, "rand", lambda *a, **kw: perm.clone())
            xg, loss = h.forward_backward(y.to(dev), x.to(dev), o)
            monkeypatch.setattr(torch, "rand", real_rand)
            if k == 0:
                o.step()
                h.update()
            else:
                h.step(o)
            o.zero_grad()
            outs.append((xg.clone(), float(loss)))
        torch.cuda.synchronize()
        a, b = heads
        assert rel(b.weight, a.weight) < 1e-6 and rel(b.weight_mom, a.weight_mom) < 1e-6, (it, rel(b.weight, a.weight))
        assert abs(outs[-1][1] - outs[-2][1]) <= 1e-5 * abs(outs[-2][1])
        assert rel(outs[-1][0], outs[-2][0]) < 1e-3
    if sample_rate == 1.0 and not nesterov and dampening == 0.0:
        assert torch.equal(heads[0].weight, heads[1].weight), "same arithmetic as torch's foreach SGD -> same bits"


# ---------------------------------------------------------------------------------------------------------------
# stored-probability backward (the default of the tensor path) against the recomputing backward and the oracle
# ---------------------------------------------------------------------------------------------------------------
def _run_mode(pkg, mode, cfg, w, x, y):
    head = _make_head(pkg, cfg, w, False)
    head._ops.bwd_mode = mode
    opt = torch.optim.SGD([{"params": head.parameters()}], lr=0.1, momentum=0.9)
    x_grad, loss = head.forward_backward(y.to(head.device), x.to(head.device), opt)
    return x_grad.cpu(), float(loss), head.sub_weight.grad.cpu()


@pytest.mark.parametrize("B,C,E,s,margin", [(512, 30000, 512, 64.0, "cosface"), (300, 5000, 256, 30.0, "arcface"), (130, 700, 128, 64.0, "cosface"),
                                             (70, 300, 64, 64.0, "cosface")])
def test_backward_modes_agree(pkg, B, C, E, s, margin):
    """'prob' (forward keeps exp2(logit - bound), no recomputation GEMM) and 'recompute' are two routes to the same
    partial_fc.py:150-168 quantities: each within the bf16 tolerance of the oracle, every embedding size / cluster shape."""
    from oracle import partial_fc_oracle as O
    g = torch.Generator().manual_seed(B + C + E)
    w = torch.randn(C, E, generator=g) * 0.01
    y = torch.randint(0, C, (B,), generator=g)
    y[1] = y[0]                                                        # two rows of one class
    x = torch.nn.functional.normalize(torch.randn(B, E, generator=g) + 2.0 * torch.nn.functional.normalize(w[y]) * (torch.arange(B) % 4 == 0)[:, None])
    m = 0.4 if margin == "cosface" else 0.5
    ref = O.forward_backward([x], [y], [w], C, s, m, margin=margin)
    cfg = dict(batch=B, num_classes=C, emb=E, s=s, m=m, sample_rate=1.0, loss=margin)
    for mode in ("prob", "recompute"):
        dx, loss, dw = _run_mode(pkg, mode, cfg, w, x, y)
        assert abs(loss - float(ref.loss)) <= 1e-2 * float(ref.loss), mode
        assert rel(dx, ref.x_grad[0]) < 1e-2, mode
        assert rel(dw, ref.dw[0]) < 1e-2, mode


def test_prob_unnormalised_features(pkg):
    """forward_backward uses the features as passed (partial_fc.py:110): row norms from 0.3 to 1.6 move the per-row
    bound of the stored-probability forward; s |x| stays inside its documented range (include/fedfr_b200.h)."""
    from oracle import partial_fc_oracle as O
    B, C, E, s = 256, 6000, 512, 30.0
    g = torch.Generator().manual_seed(77)
    w = torch.randn(C, E, generator=g) * 0.01
    y = torch.randint(0, C, (B,), generator=g)
    x = torch.nn.functional.normalize(torch.randn(B, E, generator=g)) * torch.linspace(0.3, 1.6, B)[:, None]
    x[3] = 0.0                                                         # a zero row: every logit 0, uniform softmax
    ref = O.forward_backward([x], [y], [w], C, s, 0.4)
    dx, loss, dw = _run_mode(pkg, "prob", dict(batch=B, num_classes=C, emb=E, s=s, m=0.4, sample_rate=1.0), w, x, y)
    assert abs(loss - float(ref.loss)) <= 1e-2 * float(ref.loss)
    assert rel(dx, ref.x_grad[0]) < 1e-2
    assert rel(dw, ref.dw[0]) < 1e-2


@pytest.mark.parametrize("mode", ["prob", "recompute"])
def test_hard_sample_clamp_tensor_path(pkg, mode):
    """cos = -1 to the own class with s = 64: the target probability underflows, the loss takes the 1e-30 clamp
    (partial_fc.py:162) and the target gradient is -s / Bt -- also when the probability was never stored."""
    from oracle import partial_fc_oracle as O
    B, C, E = 8, 300, 64
    g = torch.Generator().manual_seed(5)
    w = torch.randn(C, E, generator=g)
    x = torch.nn.functional.normalize(torch.randn(B, E, generator=g))
    y = torch.randint(0, C, (B,), generator=g)
    x[0] = -torch.nn.functional.normalize(w[y[0]], dim=0)
    ref = O.forward_backward([x], [y], [w], C, 64.0, 0.4)
    dx, loss, dw = _run_mode(pkg, mode, dict(batch=B, num_classes=C, emb=E, s=64.0, m=0.4, sample_rate=1.0), w, x, y)
    assert abs(loss - float(ref.loss)) < 1e-2 * float(ref.loss)
    assert rel(dx, ref.x_grad[0]) < 1e-2 and rel(dw, ref.dw[0]) < 1e-2


def test_large_norm_features_fall_back_to_recompute(pkg):
    """Un-normalised embeddings with |x| = 3 (s |x| = 192 nats) are outside the stored-probability window: the head
    must notice on its first call, switch to the recomputing backward and still match the oracle.  (The bf16 logit
    error grows with s |x|, so this case is held to 3e-2 instead of 1e-2.)"""
    from oracle import partial_fc_oracle as O
    B, C, E, s = 128, 3000, 256, 64.0
    g = torch.Generator().manual_seed(123)
    w = torch.randn(C, E, generator=g) * 0.01
    y = torch.randint(0, C, (B,), generator=g)
    x = torch.nn.functional.normalize(torch.randn(B, E, generator=g)) * 3.0
    ref = O.forward_backward([x], [y], [w], C, s, 0.4)
    head = _make_head(pkg, dict(batch=B, num_classes=C, emb=E, s=s, m=0.4, sample_rate=1.0), w, False)
    assert head._ops.bwd_mode == "prob"
    opt = torch.optim.SGD([{"params": head.parameters()}], lr=0.1, momentum=0.9)
    dx, loss = head.forward_backward(y.to(head.device), x.to(head.device), opt)
    assert head._ops.bwd_mode == "recompute"
    assert abs(float(loss) - float(ref.loss)) <= 3e-2 * float(ref.loss)
    assert rel(dx, ref.x_grad[0]) < 3e-2
    assert rel(head.sub_weight.grad, ref.dw[0]) < 3e-2
