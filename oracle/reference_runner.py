"""Run the UNMODIFIED reference ``partial_fc.PartialFC`` (+ ``losses``) -- TEST / BASELINE INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__`` and the reference / cpu_baseline legs of ``bench.py`` import this; the product
(``fedfr_b200/``) never does.  The reference's two source files are not part of this repository: ``__graft_entry__.build()``
copies them, where ``/root/reference`` exists, into the git-ignored ``baseline/_ref/`` so that they travel to the GPU box
with the snapshot; this module imports them from there (or straight from ``/root/reference``) and drives the class
through the shims of SURVEY 8c:

* CPU: the constructor hard-codes CUDA (partial_fc.py:27,61,69), so the module is assembled with ``__new__`` and the very
  same field assignments on ``device=cpu``; ``torch.cuda.current_stream`` / ``torch.cuda.stream`` become no-ops
  (partial_fc.py:109,119).
* any device: ``dist.reduce_scatter`` is called under ``torch.no_grad()`` (partial_fc.py:171-173 writes in place into a
  ``requires_grad`` leaf, which torch >= 2 refuses).

Everything else -- ``prepare``, ``sample``, ``forward``, the margin, the hand-written softmax, ``logits.backward``,
the collectives -- is the reference's own code on its stock path.
"""
import contextlib
import importlib.util
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIRS = [os.path.join(ROOT, "baseline", "_ref"), "/root/reference"]
FILES = ("partial_fc.py", "losses.py")

_mods = None


def available():
    return any(all(os.path.exists(os.path.join(d, f)) for f in FILES) for d in REF_DIRS)


def reference_dir():
    for d in REF_DIRS:
        if all(os.path.exists(os.path.join(d, f)) for f in FILES):
            return d
    raise FileNotFoundError("the reference's partial_fc.py / losses.py are neither in baseline/_ref nor in /root/reference "
                            "(run __graft_entry__.build() in the build container)")


def load():
    """(partial_fc, losses) modules of the unmodified reference, under private names so that they never shadow this
    repo's own ``losses`` / ``partial_fc`` attributes."""
    global _mods
    if _mods is None:
        d = reference_dir()
        out = []
        for name in ("partial_fc", "losses"):
            spec = importlib.util.spec_from_file_location("fedfr_reference_" + name, os.path.join(d, name + ".py"))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[spec.name] = mod
            spec.loader.exec_module(mod)
            out.append(mod)
        _mods = tuple(out)
    return _mods


@contextlib.contextmanager
def shims(cpu):
    """Process-wide patches for the duration of a reference run, restored on exit (bench.py goes back to the CUDA path)."""
    import torch
    import torch.distributed as dist
    saved = (dist.reduce_scatter, torch.cuda.current_stream, torch.cuda.stream)
    real_rs = dist.reduce_scatter

    def rs(out, ins, *a, **k):
        with torch.no_grad():
            return real_rs(out, ins, *a, **k)
    dist.reduce_scatter = rs
    if cpu:
        class _S:
            def wait_stream(self, *_):
                pass
        torch.cuda.current_stream = lambda *a, **k: _S()
        torch.cuda.stream = lambda *_a, **_k: contextlib.nullcontext()
    try:
        yield
    finally:
        dist.reduce_scatter, torch.cuda.current_stream, torch.cuda.stream = saved


def load_server_functions():
    """``FedPavg`` and ``FedAvg_on_FC`` of the unmodified reference (server.py:25-46), compiled from the staged source file
    without importing the module (its top level pulls in mxnet / easydict, which this image does not have): the two
    function definitions are taken out of the file's AST verbatim and executed with the names they use (copy, torch)."""
    import ast
    import copy
    import torch
    path = None
    for d in REF_DIRS:
        if os.path.exists(os.path.join(d, "server.py")):
            path = os.path.join(d, "server.py")
            break
    if path is None:
        raise FileNotFoundError("server.py of the reference is neither in baseline/_ref nor in /root/reference")
    tree = ast.parse(open(path).read(), filename=path)
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("FedPavg", "FedAvg_on_FC")]
    ns = {"copy": copy, "torch": torch}
    exec(compile(ast.Module(body=fns, type_ignores=[]), path, "exec"), ns)
    return ns["FedPavg"], ns["FedAvg_on_FC"]


def build_head(device, rank, world_size, batch_size, num_classes, sample_rate, emb, s, m, weight=None, loss="cosface", local_rank=0):
    """The reference module with ``weight`` (fp32 [num_local, emb]; None = its own N(0, 0.01) init) on ``device``."""
    import torch
    from torch.nn import Module
    from torch.nn.parameter import Parameter
    partial_fc, losses = load()
    margin = (losses.ArcFace if loss == "arcface" else losses.CosFace)(s=s, m=m)
    device = torch.device(device)
    if device.type == "cuda":                       # the real constructor, partial_fc.py:19-69
        mod = partial_fc.PartialFC(rank, local_rank, world_size, batch_size, False, margin, num_classes, sample_rate, emb, "/tmp")
        if weight is not None:
            with torch.no_grad():
                mod.weight.copy_(weight)
        return mod
    mod = partial_fc.PartialFC.__new__(partial_fc.PartialFC)
    Module.__init__(mod)
    mod.num_classes, mod.rank, mod.local_rank = num_classes, rank, local_rank
    mod.device = device
    mod.world_size, mod.batch_size = world_size, batch_size
    mod.margin_softmax = margin
    mod.sample_rate, mod.embedding_size, mod.prefix = sample_rate, emb, "./"
    mod.num_local = num_classes // world_size + int(rank < num_classes % world_size)
    mod.class_start = num_classes // world_size * rank + min(rank, num_classes % world_size)
    mod.num_sample = int(sample_rate * mod.num_local)
    mod.weight = torch.normal(0, 0.01, (mod.num_local, emb)) if weight is None else weight.clone()
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


@contextlib.contextmanager
def single_rank_group(backend):
    """The reference calls torch.distributed unconditionally (partial_fc.py:122-173): give it a real 1-rank group."""
    import torch.distributed as dist
    if dist.is_initialized():
        yield
        return
    if os.environ.get("TORCHELASTIC_USE_AGENT_STORE"):
        # A torchrun worker (bench.py --impl reference --gpus N>1, rank 0): the inherited switch turns every tcp:// or env://
        # rendezvous into a CLIENT of the agent's store, so a private 1-rank group on its own port would wait for a server
        # that does not exist.  An in-process store needs neither a socket nor the environment.
        dist.init_process_group(backend, store=dist.HashStore(), rank=0, world_size=1)
    else:
        import socket
        with socket.socket() as so:
            so.bind(("127.0.0.1", 0))
            port = so.getsockname()[1]
        dist.init_process_group(backend, init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1)
    try:
        yield
    finally:
        dist.destroy_process_group()


def time_steps(device, batch, num_classes, emb, sample_rate, s, m, steps, warmup, seed=100, threads=None):
    """Per-step wall times (s) of the unmodified ``forward_backward`` at world size 1 on ``device`` ("cpu" or "cuda:N"):
    bench.py's synthetic inputs, ``sub_weight.grad`` reset every step as ``zero_grad(set_to_none=True)`` does."""
    import torch
    dev = torch.device(device)
    cpu = dev.type == "cpu"
    if cpu and threads:
        torch.set_num_threads(threads)
    times = []
    with shims(cpu), single_rank_group("gloo" if cpu else "nccl"):
        g = torch.Generator().manual_seed(seed)
        feats = torch.nn.functional.normalize(torch.randn(batch, emb, generator=g)).to(dev)
        label = torch.randint(0, num_classes, (batch,), generator=g).to(dev)
        head = build_head(dev, 0, 1, batch, num_classes, sample_rate, emb, s, m)
        opt = torch.optim.SGD([{"params": head.parameters()}], lr=0.1, momentum=0.9, weight_decay=5e-4)
        for i in range(warmup + steps):
            head.sub_weight.grad = None
            if not cpu:
                torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            x_grad, loss = head.forward_backward(label, feats, opt)
            if not cpu:
                torch.cuda.synchronize(dev)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
        out = {"loss": float(loss.detach()), "x_grad_norm": float(x_grad.detach().norm())}
        del head, opt
    return times, out
