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
    """Pinned host staging for the pointer table (reused across calls)."""

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


def weighted_sum_segments(srcs: List[List[torch.Tensor]], weights_f32: List[float], device) -> List[torch.Tensor]:
    """srcs[s][i] = tensor of client i for segment s (all on ``device``, contiguous).  Returns fp32 outputs."""
    n_seg, k = len(srcs), len(weights_f32)
    _tables.ensure(n_seg, k)
    outs = []
    for s, group in enumerate(srcs):
        ref = group[0]
        if ref.dtype == torch.float32:
            code = N.FEDAVG_F32
        elif ref.dtype == torch.int64:
            code = N.FEDAVG_I64
        else:
            raise TypeError(f"FedPavg: unsupported state_dict dtype {ref.dtype} (reference models hold fp32 + int64 counters)")
        out = torch.empty(ref.shape, dtype=torch.float32, device=device)
        outs.append(out)
        _tables.out[s] = out.data_ptr()
        _tables.len[s] = ref.numel()
        _tables.dtype[s] = code
        for i, t in enumerate(group):
            if t.dtype != ref.dtype or t.shape != ref.shape or not t.is_contiguous() or t.device != out.device:
                raise ValueError("FedPavg: every client must hold the same dtype/shape (contiguous, on the reduction device) per key")
            _tables.src[s * k + i] = t.data_ptr()
    for i, w in enumerate(weights_f32):
        _tables.w[i] = w
    nbytes = N.lib.fedavg_table_bytes(n_seg, k)
    key = (device.index, nbytes)
    tab = _dev_table.get(device.index)
    if tab is None or tab.numel() < nbytes:
        tab = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _dev_table[device.index] = tab
    st = torch.cuda.current_stream(device).cuda_stream
    with torch.cuda.device(device):
        N.check(N.lib.fedavg_weighted_sum(_tables.src.data_ptr(), _tables.out.data_ptr(), _tables.len.data_ptr(), _tables.dtype.data_ptr(),
                                          n_seg, _tables.w.data_ptr(), k, tab.data_ptr(), tab.numel(), st), "fedavg_weighted_sum")
    return outs


def FedPavg(models: List[Dict[str, torch.Tensor]], weights: Sequence[float], device=None, out_device=None):
    """server.py:25-34."""
    if not torch.cuda.is_available():
        raise RuntimeError("fedfr_b200.FedPavg needs a CUDA device (sm_100); there is no CPU fallback")
    dev = _device_of(models, device)
    wn = [float(torch.tensor(w, dtype=torch.float64).to(torch.float32)) for w in _normalised_weights(weights)]
    keys = list(models[0].keys())
    srcs = []
    for name in keys:
        group = []
        for sd in models:
            t = sd[name]
            if t.device != dev:
                t = (t if t.is_pinned() or t.is_cuda else t.pin_memory()).to(dev, non_blocking=True)
            group.append(t.contiguous())
        srcs.append(group)
    outs = weighted_sum_segments(srcs, wn, dev)
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
