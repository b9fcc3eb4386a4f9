"""The oracle (oracle/partial_fc_oracle.py) against vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from golden_util import Case, SMALL_CASES
from oracle import partial_fc_oracle as O


def _run_steps(case, dtype=torch.float32):
    cfg = case.cfg
    W = cfg["world_size"]
    weights = [w.clone() for w in case.weights]
    moms = [torch.zeros_like(w) for w in weights]
    for step in range(cfg["steps"]):
        perms = [case.get(r, step, "perm") if case.has(r, step, "perm") else None for r in range(W)]
        out = O.forward_backward(case.features, case.labels, weights, cfg["num_classes"], cfg["s"], cfg["m"],
                                 cfg["sample_rate"], perms, dtype=dtype, margin=cfg.get("loss", "cosface"))
        yield step, out, weights, moms
        for r in range(W):          # optimizer.step() + update()  (partial_fc.py:113-116,124-126)
            if out.index[r] is None:
                weights[r], moms[r] = O.sgd_momentum_step(weights[r], moms[r], out.dw[r].float(), cfg["lr"])
            else:
                idx = torch.from_numpy(out.index[r])
                w_s, m_s = O.sgd_momentum_step(weights[r][idx], moms[r][idx], out.dw[r].float(), cfg["lr"])
                weights[r] = weights[r].clone(); moms[r] = moms[r].clone()
                weights[r][idx], moms[r][idx] = w_s, m_s


@pytest.mark.parametrize("name", SMALL_CASES)
def test_oracle_matches_reference(name):
    case = Case(name)
    W = case.cfg["world_size"]
    # ArcFace: the reference pushes EVERY logit through acos_/cos_ (losses.py:43-44), an identity that costs ~1e-7 of
    # rounding per element before the x s scale; the oracle only touches the target column -> a wider absolute floor
    atol = 1e-5 if case.cfg.get("loss", "cosface") == "arcface" else 2e-6
    for step, out, weights, moms in _run_steps(case):
        for r in range(W):
            assert abs(float(out.loss) - float(case.get(r, step, "loss"))) <= 2e-5 * max(1.0, abs(float(out.loss)))
            np.testing.assert_allclose(out.x_grad[r].numpy(), case.get(r, step, "x_grad"), rtol=2e-4, atol=atol)
            np.testing.assert_allclose(out.dw[r].numpy(), case.get(r, step, "dw"), rtol=2e-4, atol=atol)
            if case.has(r, step, "index"):
                np.testing.assert_array_equal(out.index[r], case.get(r, step, "index"))      # bit exact


@pytest.mark.parametrize("name", ["w1_sr1_small", "w1_sr01", "w2_sr03"])
def test_oracle_weight_update_matches_reference(name):
    case = Case(name)
    cfg = case.cfg
    W = cfg["world_size"]
    weights = [w.clone() for w in case.weights]
    moms = [torch.zeros_like(w) for w in weights]
    for step in range(cfg["steps"]):
        perms = [case.get(r, step, "perm") if case.has(r, step, "perm") else None for r in range(W)]
        out = O.forward_backward(case.features, case.labels, weights, cfg["num_classes"], cfg["s"], cfg["m"],
                                 cfg["sample_rate"], perms)
        for r in range(W):
            if out.index[r] is None:
                weights[r], moms[r] = O.sgd_momentum_step(weights[r], moms[r], out.dw[r], cfg["lr"])
            else:
                idx = torch.from_numpy(out.index[r])
                w_s, m_s = O.sgd_momentum_step(weights[r][idx], moms[r][idx], out.dw[r], cfg["lr"])
                weights[r] = weights[r].clone(); moms[r] = moms[r].clone()
                weights[r][idx], moms[r][idx] = w_s, m_s
            np.testing.assert_allclose(weights[r].numpy(), case.get(r, step, "weight_after"), rtol=3e-4, atol=3e-6)
            np.testing.assert_allclose(moms[r].numpy(), case.get(r, step, "mom_after"), rtol=3e-4, atol=3e-6)


def test_oracle_c1_config():
    """BASELINE.json configs[0]: B=128, C=10k, E=512, sr=1, W=1 -- inputs regenerated from the seed."""
    case = Case("c1_b128_c10k")
    out = O.forward_backward(case.features, case.labels, case.weights, 10000, 64.0, 0.4)
    assert abs(float(out.loss) - float(case.get(0, 0, "loss"))) < 1e-3
    np.testing.assert_allclose(out.x_grad[0].numpy(), case.get(0, 0, "x_grad"), rtol=5e-4, atol=5e-6)
    np.testing.assert_allclose(out.dw[0].numpy()[::97], case.get(0, 0, "dw_rows"), rtol=5e-4, atol=5e-6)
    assert abs(np.linalg.norm(out.dw[0].double().numpy()) / float(case.get(0, 0, "dw_norm")) - 1) < 1e-4


def test_oracle_equals_dense_twin():
    """client.py:69-74 + losses.py:23-29 + cross_entropy == PartialFC maths at W=1, sr=1."""
    case = Case("w1_sr1_small")
    out = O.forward_backward(case.features, case.labels, case.weights, case.cfg["num_classes"], 64.0, 0.4, dtype=torch.float64)
    loss, dx, dw = O.dense_twin_grads(case.features[0].double(), case.weights[0].double(), case.labels[0], 64.0, 0.4)
    assert abs(float(loss) - float(out.loss)) < 1e-9
    np.testing.assert_allclose(out.x_grad[0].numpy(), dx.numpy(), rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(out.dw[0].numpy(), dw.numpy(), rtol=1e-8, atol=1e-12)


def test_multi_rank_equals_single_rank():
    """W ranks reproduce the W=1 result on the concatenated batch (SURVEY 8c identities)."""
    case = Case("w2_sr1_ragged")
    cfg = case.cfg
    out2 = O.forward_backward(case.features, case.labels, case.weights, cfg["num_classes"], dtype=torch.float64)
    out1 = O.forward_backward([torch.cat(case.features)], [torch.cat(case.labels)], [torch.cat(case.weights)],
                              cfg["num_classes"], dtype=torch.float64)
    assert abs(float(out1.loss) - float(out2.loss)) < 1e-10
    np.testing.assert_allclose(torch.cat(out2.x_grad).numpy(), 2 * out1.x_grad[0].numpy(), rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(torch.cat(out2.dw).numpy(), out1.dw[0].numpy(), rtol=1e-9, atol=1e-13)


def test_fedavg_oracle_matches_reference():
    z = np.load(__import__("os").path.join(__import__("golden_util").GOLDEN, "fedavg.npz"))
    K = int(z["K"])
    keys = sorted({k.split("/", 1)[1] for k in z.files if k.startswith("in0/")})
    models = [{k: z[f"in{i}/{k}"] for k in keys} for i in range(K)]
    out = O.fedpavg(models, [int(w) for w in z["weights"]])
    for k in keys:
        ref = z[f"out/{k}"]
        assert out[k].dtype == np.float32 and ref.dtype == np.float32
        np.testing.assert_array_equal(out[k], ref)          # bit exact, incl. the int64 -> fp32 counter
    fcs = [z[f"fc_in{i}"] for i in range(K)]
    w = [int(v) for v in z["weights"]]
    np.testing.assert_array_equal(O.fedavg_on_fc(z["fc_old"], fcs, w, 1), z["fc_out_p1"])
    np.testing.assert_array_equal(O.fedavg_on_fc(z["fc_old"], fcs, w, 0.7), z["fc_out_p07"])


def test_sampling_invariants():
    rng = np.random.default_rng(0)
    for _ in range(50):
        nl = int(rng.integers(20, 400)); k = int(rng.integers(0, nl))
        y = rng.integers(-1, nl, size=64)
        perm = rng.random(nl, dtype=np.float32)
        idx = O.sample_index(y, perm, k)
        pos = np.unique(y[y >= 0])
        assert np.all(np.diff(idx) > 0) and np.isin(pos, idx).all()
        assert idx.size == max(k, pos.size)
        ry = O.relabel_to_sample(y, idx)
        assert np.array_equal(idx[ry[y >= 0]], y[y >= 0]) and np.all(ry[y < 0] == -1)
