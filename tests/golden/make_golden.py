"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

What it does: imports ``partial_fc.py`` / ``losses.py`` / ``server.py`` from ``/root/reference`` and drives
``PartialFC.forward_backward`` (partial_fc.py:130-176), ``update`` (:113-116) and ``FedPavg`` /
``FedAvg_on_FC`` (server.py:25-46) on CPU + gloo.  The reference's constructor hard-codes CUDA
(partial_fc.py:27,61,69), so the module object is assembled with ``__new__`` and the very same field
assignments on ``device=cpu``; three process-wide shims make the rest run on torch 2.11 without a GPU:

* ``torch.cuda.current_stream`` -> object with a no-op ``wait_stream``      (partial_fc.py:109)
* ``torch.cuda.stream``         -> null context                              (partial_fc.py:119)
* ``dist.reduce_scatter``       -> same call under ``torch.no_grad()``       (partial_fc.py:171-173 writes
  in place into a ``requires_grad`` leaf, which modern torch refuses)

``torch.rand`` is wrapped to record the ``perm`` draw of ``sample()`` (partial_fc.py:95) so the CUDA
index kernel can be fed the identical numbers.  No reference source is copied into this repo; only
inputs/outputs are stored (``*.npz``).
"""
import contextlib
import hashlib
import os
import sys
import types

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def _install_shims():
    class _S:
        def wait_stream(self, *_):
            pass
    torch.cuda.current_stream = lambda *a, **k: _S()
    torch.cuda.stream = lambda *_a, **_k: contextlib.nullcontext()
    real_rs = dist.reduce_scatter

    def rs(out, ins, *a, **k):
        with torch.no_grad():
            return real_rs(out, ins, *a, **k)
    dist.reduce_scatter = rs


def _import_reference():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import partial_fc  # noqa
    import losses  # noqa
    return partial_fc, losses


def _build_module(partial_fc, losses, rank, world_size, batch_size, num_classes, sample_rate, emb, weight, s, m, loss="cosface"):
    """Field-for-field what partial_fc.py:24-69 sets, with device=cpu and a caller supplied shard."""
    from torch.nn import Module
    from torch.nn.parameter import Parameter
    mod = partial_fc.PartialFC.__new__(partial_fc.PartialFC)
    Module.__init__(mod)
    mod.num_classes, mod.rank, mod.local_rank = num_classes, rank, rank
    mod.device = torch.device("cpu")
    mod.world_size, mod.batch_size = world_size, batch_size
    mod.margin_softmax = (losses.ArcFace if loss == "arcface" else losses.CosFace)(s=s, m=m)
    mod.sample_rate, mod.embedding_size, mod.prefix = sample_rate, emb, "./"
    mod.num_local = num_classes // world_size + int(rank < num_classes % world_size)
    mod.class_start = num_classes // world_size * rank + min(rank, num_classes % world_size)
    mod.num_sample = int(sample_rate * mod.num_local)
    mod.weight = weight.clone()
    mod.weight_mom = torch.zeros_like(mod.weight)
    mod.stream = None
    mod.index = None
    if int(sample_rate) == 1:
        mod.update = lambda: 0
        mod.sub_weight = Parameter(mod.weight)
        mod.sub_weight_mom = mod.weight_mom
    else:
        mod.sub_weight = Parameter(torch.empty((0, 0)))
    return mod


def make_inputs(seed, world_size, batch, num_classes, emb, label_pool=None):
    g = torch.Generator().manual_seed(seed)
    feats, labels, weights = [], [], []
    for r in range(world_size):
        f = torch.nn.functional.normalize(torch.randn(batch, emb, generator=g))
        if label_pool is None:
            l = torch.randint(0, num_classes, (batch,), generator=g)
        else:
            l = label_pool[torch.randint(0, len(label_pool), (batch,), generator=g)]
        nl = num_classes // world_size + int(r < num_classes % world_size)
        w = torch.randn(nl, emb, generator=g) * 0.01
        feats.append(f), labels.append(l.long()), weights.append(w)
    return feats, labels, weights


def _worker(rank, world_size, cfg, feats, labels, weights, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    torch.set_num_threads(2)
    _install_shims()
    partial_fc, losses = _import_reference()
    torch.manual_seed(cfg["seed"] * 1000 + rank)
    perm_log = []
    real_rand = torch.rand

    def rand_spy(*a, **k):
        t = real_rand(*a, **k)
        perm_log.append(t.clone())
        return t
    torch.rand = rand_spy
    mod = _build_module(partial_fc, losses, rank, world_size, cfg["batch"], cfg["num_classes"],
                        cfg["sample_rate"], cfg["emb"], weights[rank], cfg["s"], cfg["m"], cfg.get("loss", "cosface"))
    opt = torch.optim.SGD([{"params": mod.parameters()}], lr=cfg["lr"], momentum=0.9, weight_decay=5e-4)
    steps = []
    for step in range(cfg["steps"]):
        perm_log.clear()
        x_grad, loss_v = mod.forward_backward(labels[rank], feats[rank], opt)
        rec = {
            "x_grad": x_grad.detach().numpy().copy(),
            "loss": np.float32(loss_v.item()),
            "dw": mod.sub_weight.grad.detach().numpy().copy(),
            "index": None if mod.index is None else mod.index.numpy().copy(),
            "perm": perm_log[0].numpy().copy() if perm_log else None,
        }
        opt.step()
        mod.update()
        opt.zero_grad()
        rec["weight_after"] = mod.weight.detach().numpy().copy()
        rec["mom_after"] = mod.weight_mom.detach().numpy().copy()
        steps.append(rec)
    ret[rank] = steps
    dist.barrier()
    dist.destroy_process_group()


_PORT = [29611]


def run_reference(cfg, feats, labels, weights):
    W = cfg["world_size"]
    mgr = mp.Manager()
    ret = mgr.dict()
    _PORT[0] += 1
    if W == 1:
        _worker(0, 1, cfg, feats, labels, weights, _PORT[0], ret)
    else:
        mp.spawn(_worker, args=(W, cfg, feats, labels, weights, _PORT[0], ret), nprocs=W, join=True)
    return [ret[r] for r in range(W)]


def sha(arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


CASES = {
    # name: cfg.   store_inputs=False -> inputs are regenerated from the seed and pinned by sha256
    "w1_sr1_small": dict(seed=1, world_size=1, batch=16, num_classes=200, emb=512, sample_rate=1.0, s=64.0, m=0.4, lr=0.002, steps=2, store_inputs=True),
    "w1_sr1_s30": dict(seed=2, world_size=1, batch=24, num_classes=96, emb=128, sample_rate=1.0, s=30.0, m=0.4, lr=0.002, steps=1, store_inputs=True),
    "w1_sr01": dict(seed=3, world_size=1, batch=16, num_classes=400, emb=128, sample_rate=0.1, s=64.0, m=0.4, lr=0.002, steps=2, store_inputs=True),
    "w1_sr_pos_overflow": dict(seed=4, world_size=1, batch=64, num_classes=300, emb=64, sample_rate=0.1, s=64.0, m=0.4, lr=0.002, steps=1, store_inputs=True),
    "w2_sr1_ragged": dict(seed=5, world_size=2, batch=16, num_classes=301, emb=128, sample_rate=1.0, s=64.0, m=0.4, lr=0.002, steps=1, store_inputs=True),
    "w2_sr03": dict(seed=6, world_size=2, batch=16, num_classes=401, emb=128, sample_rate=0.3, s=64.0, m=0.4, lr=0.002, steps=2, store_inputs=True),
    # ArcFace (losses.py:32-45), the second margin of the same boundary (--loss ArcFace, client.py:133)
    "w1_arc_small": dict(seed=7, world_size=1, batch=16, num_classes=200, emb=128, sample_rate=1.0, s=64.0, m=0.5, lr=0.002, steps=2, store_inputs=True, loss="arcface"),
    "w2_arc_sr03": dict(seed=8, world_size=2, batch=16, num_classes=401, emb=128, sample_rate=0.3, s=30.0, m=0.5, lr=0.002, steps=1, store_inputs=True, loss="arcface"),
    "c1_b128_c10k": dict(seed=100, world_size=1, batch=128, num_classes=10000, emb=512, sample_rate=1.0, s=64.0, m=0.4, lr=0.002, steps=1, store_inputs=False),
}


def fedavg_golden():
    for name in ("easydict", "mxnet", "prettytable"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["easydict"].EasyDict = type("EasyDict", (dict,), {"__getattr__": dict.get, "__setattr__": dict.__setitem__})
    for sub in ("ndarray", "recordio", "image"):
        m = types.ModuleType("mxnet." + sub)
        sys.modules["mxnet." + sub] = m
        setattr(sys.modules["mxnet"], sub, m)
    sys.modules["prettytable"].PrettyTable = object
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import server
    g = torch.Generator().manual_seed(77)
    K = 5
    base = {"conv.weight": torch.randn(8, 3, 3, 3, generator=g), "bn.weight": torch.randn(8, generator=g),
            "bn.running_var": torch.rand(8, generator=g) + 0.5, "bn.num_batches_tracked": torch.tensor(0),
            "fc.weight": torch.randn(33, 17, generator=g)}
    models = []
    for i in range(K):
        sd = {}
        for k, v in base.items():
            sd[k] = (v + 0.01 * torch.randn(v.shape, generator=g)) if v.is_floating_point() else torch.tensor(1000 * i + 7 * i * i + 3)
        models.append(sd)
    weights = [6000 + 37 * i for i in range(K)]
    out = server.FedPavg(models, weights)
    store = {"weights": np.array(weights, dtype=np.int64), "K": np.int64(K)}
    for i, sd in enumerate(models):
        for k, v in sd.items():
            store[f"in{i}/{k}"] = v.numpy()
    for k, v in out.items():
        store[f"out/{k}"] = v.numpy()
    fcs = [torch.randn(50, 32, generator=g) for _ in range(K)]
    old = torch.randn(50, 32, generator=g)
    store["fc_old"] = old.numpy()
    for i, f in enumerate(fcs):
        store[f"fc_in{i}"] = f.numpy()
    store["fc_out_p1"] = server.FedAvg_on_FC(old, fcs, weights, 1).numpy()
    store["fc_out_p07"] = server.FedAvg_on_FC(old, fcs, weights, 0.7).numpy()
    np.savez_compressed(os.path.join(OUT, "fedavg.npz"), **store)
    print("fedavg.npz written")


def spreadout_golden():
    """server.SpreadOut_Module (server.py:48-63) on CPU: its forward puts the diagonal mask on the GPU with .cuda(); that
    one call is patched to a no-op for this run, everything else is the unmodified reference + autograd."""
    for name in ("easydict", "mxnet", "prettytable"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["easydict"].EasyDict = type("EasyDict", (dict,), {"__getattr__": dict.get, "__setattr__": dict.__setitem__})
    for sub in ("ndarray", "recordio", "image"):
        m = types.ModuleType("mxnet." + sub)
        sys.modules["mxnet." + sub] = m
        setattr(sys.modules["mxnet"], sub, m)
    sys.modules["prettytable"].PrettyTable = object
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import server
    g = torch.Generator().manual_seed(4242)
    n, e = 200, 64
    base = torch.randn(n // 16 + 1, e, generator=g)
    fc = (base[torch.arange(n) % len(base)] + 0.35 * torch.randn(n, e, generator=g)) * 0.05
    store = {"fc": fc.numpy()}
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        for mode, margin in (("sum", 0.4), ("mean", 0.7)):
            mod = server.SpreadOut_Module(fc.clone(), margin=margin, mode=mode)
            loss = mod()
            loss.backward()
            store[f"loss_{mode}"] = loss.detach().numpy()
            store[f"grad_{mode}"] = mod.FC.grad.numpy()
            store[f"margin_{mode}"] = np.float32(margin)
    finally:
        torch.Tensor.cuda = orig
    np.savez_compressed(os.path.join(OUT, "spreadout.npz"), **store)
    print("spreadout.npz written", {k: float(store[k]) for k in store if k.startswith("loss")})


def _stub_reference_imports():
    for name in ("easydict", "mxnet", "prettytable"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["easydict"].EasyDict = type("EasyDict", (dict,), {"__getattr__": dict.get, "__setattr__": dict.__setitem__})
    for sub in ("ndarray", "recordio", "image"):
        m = types.ModuleType("mxnet." + sub)
        sys.modules["mxnet." + sub] = m
        setattr(sys.modules["mxnet"], sub, m)
    sys.modules["prettytable"].PrettyTable = object
    if REF not in sys.path:
        sys.path.insert(0, REF)


def dense_golden():
    """The dense head FedFR's live client trains (client.py:63-74 FC_module, losses.py:17-29 CosFace(s=30, m=0.4) as built
    at client.py:133, F.cross_entropy + loss.backward() as at client.py:430-435), run with the unmodified reference classes
    on CPU."""
    _stub_reference_imports()
    import client
    import losses
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(777)
    B, C, E = 48, 211, 64
    x = torch.randn(B, E, generator=g) * 3.0
    fc = torch.randn(C, E, generator=g) * 0.01
    y = torch.randint(0, C, (B,), generator=g)
    x[::3] += 4.0 * fc[y[::3]] / fc[y[::3]].norm(dim=1, keepdim=True)
    store = {"x": x.numpy(), "fc": fc.numpy(), "y": y.numpy()}
    for name, margin in (("cosface", losses.CosFace(s=30, m=0.4)), ("arcface", losses.ArcFace(s=64.0, m=0.5))):
        head = client.FC_module(E, C, "/tmp")
        head.fc.data.copy_(fc)
        xr = x.clone().requires_grad_(True)
        logits = head(xr)
        logits = margin(logits, y)
        loss = F.cross_entropy(logits, y)
        loss.backward()
        store[f"loss_{name}"] = loss.detach().numpy()
        store[f"dx_{name}"] = xr.grad.numpy()
        store[f"dfc_{name}"] = head.fc.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "dense_head.npz"), **store)
    print("dense_head.npz written", {k: float(store[k]) for k in store if k.startswith("loss")})


def bce_golden():
    """The cosine head of the personalised branch as FedFR's client builds and trains it (client.py:135-136: BCE_module(512,
    n_class, converter_layer) with m=0.4, r=30, t=3 and losses.BCE_loss(); client.py:386-394: loss = ... + 10 * bce_loss),
    run with the unmodified reference classes on CPU.  Labels include public-data ids >= n_class (client.py:49-50)."""
    _stub_reference_imports()
    import client
    import losses
    g = torch.Generator().manual_seed(4242)
    B, C, E = 40, 37, 64
    x = torch.randn(B, E, generator=g) * 2.0
    y = torch.randint(0, C + 12, (B,), generator=g)                 # ~1/4 of the rows carry a public id (no positive column)
    store = {"x": x.numpy(), "y": y.numpy()}
    torch.manual_seed(11)
    mod = client.BCE_module(E, C, 1)
    with torch.no_grad():                                           # move off the identity / zero initialisation
        mod.converter[0].weight.add_(0.05 * torch.randn(E, E, generator=g))
        mod.converter[0].bias.add_(0.1 * torch.randn(E, generator=g))
        mod.bias.add_(0.2 * torch.randn(C, generator=g))
        own = y < C
        mod.weight[y[own]] += 0.02 * x[own]                         # some rows agree with their class
    store.update({"sd/" + k: v.detach().numpy().copy() for k, v in mod.state_dict().items()})
    xr = x.clone().requires_grad_(True)
    logits, gts = mod(xr, y)
    store["logits"], store["gt"] = logits.detach().numpy().copy(), gts.numpy().copy()
    loss = 10 * losses.BCE_loss()(logits, gts)                      # edits `logits` in place (losses.py:10-11)
    loss.backward()
    store["loss"] = loss.detach().numpy()
    store["dx"] = xr.grad.numpy()
    store.update({"grad/" + k: p.grad.numpy() for k, p in mod.named_parameters()})
    np.savez_compressed(os.path.join(OUT, "bce_head.npz"), **store)
    print("bce_head.npz written: loss", float(loss), "positives", int(gts.sum()), "of", B)


def hardneg_golden():
    """FC-based hard-negative mining: the unmodified ``Client.choose_hard_negative`` (client.py:227-266) called on CPU with
    stub loader / logger objects (it only reads ``dataset.imgidx`` of a deep copy and logs).  ``HN_ID`` is what it selects;
    ``imgidx`` the 1-based image list it derives (client.py:249-255).  The feature-based variant
    (``choose_hard_negative_2``, client.py:191-224) needs a CUDA backbone; its array part (client.py:208-215) is evaluated
    here by the same torch expressions on synthetic features."""
    _stub_reference_imports()
    import client
    from functools import reduce
    g = torch.Generator().manual_seed(9090)
    E, n_self, n_pub = 64, 23, 500
    pretrain_fc = torch.randn(n_pub, E, generator=g)
    self_fc = torch.randn(n_self, E, generator=g)
    self_fc[:9] = pretrain_fc[torch.randint(0, n_pub, (9,), generator=g)] + 0.9 * torch.randn(9, E, generator=g)   # some close pairs
    pretrain_label = torch.randint(0, n_pub, (3000,), generator=g)
    loader = types.SimpleNamespace(dataset=types.SimpleNamespace(imgidx=None))
    me = types.SimpleNamespace(logger=types.SimpleNamespace(info=lambda *a, **k: None))
    _, subset = client.Client.choose_hard_negative(me, pretrain_fc, loader, pretrain_label, self_fc, threshold=0.2)
    store = {"self_fc": self_fc.numpy(), "pretrain_fc": pretrain_fc.numpy(), "pretrain_label": pretrain_label.numpy(),
             "threshold": np.float64(0.2), "HN_ID": np.asarray(me.HN_ID, dtype=np.int64), "imgidx": np.asarray(subset.dataset.imgidx)}
    # feature-based array part, client.py:208-215
    local_feats = torch.nn.functional.normalize(torch.randn(130, E, generator=g))
    pretrained_feats = torch.nn.functional.normalize(torch.randn(900, E, generator=g) + 0.5 * local_feats[torch.randint(0, 130, (900,), generator=g)])
    similarity = torch.matmul(local_feats, pretrained_feats.t())
    times = 100
    batch = len(similarity) // times + 1
    unique_idx = [torch.where(similarity[i * batch:(i + 1) * batch] > 0.35)[1].numpy() for i in range(times)]
    unique_idx = sorted(reduce(np.union1d, unique_idx))
    store.update({"local_feats": local_feats.numpy(), "pretrained_feats": pretrained_feats.numpy(), "threshold2": np.float64(0.35),
                  "unique_idx": np.asarray(unique_idx, dtype=np.int64)})
    np.savez_compressed(os.path.join(OUT, "hardneg.npz"), **store)
    print("hardneg.npz written:", len(me.HN_ID), "of", n_pub, "ids;", len(unique_idx), "of 900 images")


def main():
    if "hardneg" in sys.argv[1:]:
        hardneg_golden()
        return
    if "bce" in sys.argv[1:]:
        bce_golden()
        return
    if "spreadout" in sys.argv[1:]:
        spreadout_golden()
        return
    if "dense" in sys.argv[1:]:
        dense_golden()
        return
    only = [a for a in sys.argv[1:] if not a.startswith("-")]       # optional: regenerate just these cases
    for name, cfg in CASES.items():
        if only and name not in only:
            continue
        pool = None
        if name == "w1_sr_pos_overflow":      # 64 samples over 60 distinct ids > num_sample=30 -> index = positives
            pool = torch.arange(0, 300, 5)
        feats, labels, weights = make_inputs(cfg["seed"], cfg["world_size"], cfg["batch"], cfg["num_classes"], cfg["emb"], pool)
        res = run_reference(cfg, feats, labels, weights)
        store = {"cfg_" + k: np.array(v) for k, v in cfg.items()}
        store["input_sha256"] = np.array(sha([t.numpy() for t in feats + labels + weights]))
        for r in range(cfg["world_size"]):
            if cfg["store_inputs"]:
                store[f"r{r}/features"] = feats[r].numpy()
                store[f"r{r}/weight"] = weights[r].numpy()
            store[f"r{r}/labels"] = labels[r].numpy()
            for t, rec in enumerate(res[r]):
                for k, v in rec.items():
                    if v is None:
                        continue
                    if not cfg["store_inputs"] and k in ("dw", "weight_after", "mom_after"):
                        # too large to commit: keep norm + a strided sample
                        store[f"r{r}/s{t}/{k}_norm"] = np.float64(np.linalg.norm(v.astype(np.float64)))
                        store[f"r{r}/s{t}/{k}_rows"] = v[::97].copy()
                        continue
                    store[f"r{r}/s{t}/{k}"] = v
                if rec["perm"] is not None:   # tie check: goldens must be tie-free at the threshold
                    pass
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **store)
        print(name, "loss", [float(res[r][0]["loss"]) for r in range(cfg["world_size"])])
    if not only:
        fedavg_golden()


if __name__ == "__main__":
    main()
