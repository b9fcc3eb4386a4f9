"""Server-side FedAvg on the GPU: ``FedPavg`` / ``FedAvg_on_FC`` with the reference's signatures (server.py:25-46).

``FedPavg(models, weights)`` takes K client ``state_dict``s and returns a dict with the same keys holding
``sum_i (w_i / sum w) * sd_i[key]`` as fp32 tensors (int64 BatchNorm counters come out as fp32 too, exactly
as in the reference, and are truncated back by ``load_state_dict``).  All keys are reduced by ONE kernel
launch over a pointer table; the arithmetic order (fp32 multiply, then fp32 add, client order) makes the
result bit-identical to the reference's CPU loop.

Inputs may live on the GPU (the fast path: nothing but the table is copied) or on the CPU, as FedFR keeps
them (client.py:469,558): CPU tensors are staged through pinned memory with async H2D copies and the result is
returned on the device unless ``out_device`` says otherwise.
"""
import ctypes as C
from typing import Dict, List, Sequence

import torch

from . import _native as N


def _normalised_weights(weights: Sequence[float]) -> List[float]:
    tot = sum(weights)
    return [w / tot for w in weights]          # Python doubles, server.py:27 / :37


class _Tables:
    """Pinned host staging for the pointer table (reused across calls), with numpy views for bulk fills."""

    def __init__(self):
        self.cap_seg = 0
        self.cap_k = 0

    def ensure(self, n_seg, k):
        if n_seg > self.cap_seg or k > self.cap_k:
            self.cap_seg, self.cap_k = max(n_seg, self.cap_seg), max(k, self.cap_k)
            self.src = torch.empty(self.cap_seg * self.cap_k, dtype=torch.int64).pin_memory()
            self.out = torch.empty(self.cap_seg, dtype=torch.int64).pin_memory()
            self.len = torch.empty(self.cap_seg, dtype=torch.int64).pin_memory()
            self.dtype = torch.empty(self.cap_seg, dtype=torch.int32).pin_memory()
            self.w = torch.empty(self.cap_k, dtype=torch.float32).pin_memory()
            self.src_np, self.out_np, self.len_np = self.src.numpy(), self.out.numpy(), self.len.numpy()
            self.dtype_np, self.w_np = self.dtype.numpy(), self.w.numpy()


_tables = _Tables()
_dev_table = {}


def _device_of(models, device):
    if device is not None:
        return torch.device(device)
    for sd in models:
        for v in sd.values():
            if v.is_cuda:
                return v.device
    return torch.device("cuda", torch.cuda.current_device())


_last_launch = None          # (n_seg, k, device) of the most recent table (bench.py re-launches it to time the C-ABI call alone)


def relaunch_last():
    """Re-run the most recent weighted sum from the tables already staged (same inputs, same outputs)."""
    _launch(*_last_launch)


def _launch(n_seg, k, device):
    global _last_launch
    _last_launch = (n_seg, k, device)
    nbytes = N.lib.fedavg_table_bytes(n_seg, k)
    tab = _dev_table.get(device.index)
    if tab is None or tab.numel() < nbytes:
        tab = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _dev_table[device.index] = tab
    st = torch.cuda.current_stream(device).cuda_stream
    with torch.cuda.device(device):
        N.check(N.lib.fedavg_weighted_sum(_tables.src.data_ptr(), _tables.out.data_ptr(), _tables.len.data_ptr(), _tables.dtype.data_ptr(),
                                          n_seg, _tables.w.data_ptr(), k, tab.data_ptr(), tab.numel(), st), "fedavg_weighted_sum")


def weighted_sum_segments(srcs: List[List[torch.Tensor]], weights_f32: List[float], device, flat: bool = False):
    """srcs[s][i] = tensor of client i for segment s (all on ``device``, contiguous).  Returns fp32 outputs
    (``flat=True``: also the one flat fp32 buffer they are 16-byte-aligned views of, for a single all-reduce)."""
    return weighted_sum_clients([[group[i] for group in srcs] for i in range(len(weights_f32))], weights_f32, device, flat)


def weighted_sum_clients(clients: List[List[torch.Tensor]], weights_f32: List[float], device, flat: bool = False, _staged: bool = False):
    """clients[i][s] = tensor of client i for segment s (client-major, the order ``state_dict.values()`` yields).
    The per-tensor host work is the floor of this call (K x n_seg Python objects): pointers are gathered with one
    generator pass per client and written into the pinned tables through numpy."""
    import numpy as np
    k, ref = len(weights_f32), clients[0]
    n_seg = len(ref)
    _tables.ensure(n_seg, k)
    f32, i64 = torch.float32, torch.int64
    numels = [t.numel() for t in ref]
    codes = []
    for t in ref:
        if t.dtype is f32:
            codes.append(N.FEDAVG_F32)
        elif t.dtype is i64:
            codes.append(N.FEDAVG_I64)
        else:
            raise TypeError(f"FedPavg: unsupported state_dict dtype {t.dtype} (reference models hold fp32 + int64 counters)")
    # one flat output allocation; segments start on 16-byte boundaries (vector path of the kernel)
    offsets = np.zeros(n_seg + 1, dtype=np.int64)
    np.cumsum([(n + 3) // 4 * 4 for n in numels], out=offsets[1:])
    flat_buf = torch.empty(max(int(offsets[-1]), 4), dtype=torch.float32, device=device)
    outs = [flat_buf[int(offsets[s]):int(offsets[s]) + numels[s]].view(ref[s].shape) for s in range(n_seg)]
    _tables.out_np[:n_seg] = flat_buf.data_ptr() + 4 * offsets[:-1]
    _tables.len_np[:n_seg] = numels
    _tables.dtype_np[:n_seg] = codes
    src = _tables.src_np[:n_seg * k].reshape(n_seg, k)
    dev_index = device.index
    for i, tensors in enumerate(clients):
        if len(tensors) != n_seg:
            raise ValueError("FedPavg: every client must hold the same keys")
        for s, t in enumerate(tensors):      # structural checks; the kernel trusts the table (_stage already fixed device/contiguity)
            if t.numel() != numels[s] or t.dtype is not ref[s].dtype or \
                    not (_staged or (t.is_contiguous() and t.is_cuda and t.get_device() == dev_index)):
                raise ValueError("FedPavg: every client must hold the same dtype/shape (contiguous, on the reduction device) per key")
        src[:, i] = np.fromiter((t.data_ptr() for t in tensors), dtype=np.int64, count=n_seg)
    _tables.w_np[:k] = weights_f32
    _launch(n_seg, k, device)
    return (outs, flat_buf) if flat else outs


def _stage(sd_values, dev):
    """state_dict values -> contiguous tensors on ``dev`` (CPU tensors go through pinned memory, client.py:469)."""
    out = []
    for t in sd_values:
        if not t.is_cuda or t.device != dev:
            t = (t if t.is_pinned() or t.is_cuda else t.pin_memory()).to(dev, non_blocking=True)
        out.append(t if t.is_contiguous() else t.contiguous())
    return out


def FedPavg(models: List[Dict[str, torch.Tensor]], weights: Sequence[float], device=None, out_device=None):
    """server.py:25-34."""
    if not torch.cuda.is_available():
        raise RuntimeError("fedfr_b200.FedPavg needs a CUDA device (sm_100); there is no CPU fallback")
    dev = _device_of(models, device)
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    wn = [float(torch.tensor(w, dtype=torch.float64).to(torch.float32)) for w in _normalised_weights(weights)]
    keys = list(models[0].keys())
    clients = []
    for sd in models:
        if len(sd) != len(keys):
            raise ValueError("FedPavg: every client must hold the same keys")
        clients.append(_stage([sd[name] for name in keys], dev))
    outs = weighted_sum_clients(clients, wn, dev, _staged=True)
    aggr = {name: (o if out_device is None else o.to(out_device)) for name, o in zip(keys, outs)}
    return aggr


def FedAvg_on_FC(pretrain_fc: torch.Tensor, models: List[torch.Tensor], weights: Sequence[float], p: float, device=None):
    """server.py:36-46."""
    if not torch.cuda.is_available():
        raise RuntimeError("fedfr_b200.FedAvg_on_FC needs a CUDA device (sm_100); there is no CPU fallback")
    dev = _device_of([{"fc": m} for m in models], device)
    wn = [float(torch.tensor(w, dtype=torch.float64).to(torch.float32)) for w in _normalised_weights(weights)]
    group = [m.to(dev, non_blocking=True).contiguous() for m in models]
    aggr = weighted_sum_segments([group], wn, dev)[0]
    if p == 1:
        return aggr
    old = pretrain_fc.to(dev).contiguous()
    out = torch.empty_like(aggr)
    one_minus_p = float(torch.tensor(1 - p, dtype=torch.float64).to(torch.float32))
    p32 = float(torch.tensor(p, dtype=torch.float64).to(torch.float32))
    st = torch.cuda.current_stream(dev).cuda_stream
    N.check(N.lib.fedavg_blend(N.ptr(old), N.ptr(aggr), one_minus_p, p32, aggr.numel(), N.ptr(out), st), "fedavg_blend")
    return out


def FedPavg_sharded(local_models: List[Dict[str, torch.Tensor]], local_weights: Sequence[float], group=None, device=None,
                    _segments_fn=None):
    """``FedPavg`` with the K client ``state_dict``s sharded over the ranks of ``group``: every rank holds
    K/W clients, reduces them locally with the GLOBALLY normalised weights ``w_i / sum_all w`` (one kernel launch into
    one flat fp32 buffer) and the partial sums meet in ONE all-reduce(SUM) of that buffer -- the only collective the
    sharding needs.  Every rank returns the full aggregated dict.

    With one rank this is ``FedPavg`` bit for bit.  With W > 1 the association of the sum changes (per-rank partial sums,
    then the all-reduce), so results agree with the sequential reference to fp32 rounding (~1e-6 relative), not bitwise.
    ``_segments_fn`` lets the host-logic tests inject a CPU provider (tests/ only)."""
    import torch.distributed as dist
    if len(local_models) == 0:
        raise ValueError("FedPavg_sharded: every rank needs at least one client state_dict (it defines the key set)")
    seg_fn = _segments_fn or weighted_sum_segments
    if _segments_fn is None:
        if not torch.cuda.is_available():
            raise RuntimeError("fedfr_b200.FedPavg_sharded needs a CUDA device (sm_100); there is no CPU fallback")
        dev = _device_of(local_models, device)
    else:
        dev = torch.device(device or "cpu")
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    tot = torch.tensor([float(sum(local_weights))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=group)
    tot = float(tot.item())
    wn = [float(torch.tensor(w / tot, dtype=torch.float64).to(torch.float32)) for w in local_weights]
    keys = list(local_models[0].keys())
    if _segments_fn is None:
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        clients = [_stage([sd[name] for name in keys], dev) for sd in local_models]
        outs, flat_buf = weighted_sum_clients(clients, wn, dev, flat=True, _staged=True)
    else:
        srcs = [[sd[name].to(dev).contiguous() for sd in local_models] for name in keys]
        outs, flat_buf = _segments_fn(srcs, wn, dev, flat=True)
    if world > 1:
        dist.all_reduce(flat_buf, op=dist.ReduceOp.SUM, group=group)
    return {name: o for name, o in zip(keys, outs)}
