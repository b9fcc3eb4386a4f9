"""``BCE_module`` -- the cosine head of FedFR's personalised branch (client.py:25-60) on two fused sm_100a launches.

Same constructor, parameters (``converter.0.weight/bias``, ``weight``, ``bias`` -- so ``bce_module.pth``, client.py:485,
loads either way), ``forward(x, labels) -> (logits, gt)`` and ``initialize(fc)`` as the reference class.  The converter
(an identity-initialised ``nn.Linear``, client.py:30-34) stays a torch module; everything after it -- both
normalisations, the cosine matmul, ``g_func``, the +/- margin, ``r`` and the bias, and the matching backward -- is
``pfc_bce_head_fwd`` / ``pfc_bce_head_bwd`` (``csrc/bce_head.cu``).  The returned logits are an ordinary autograd tensor:
``losses.BCE_loss`` (losses.py:4-15), which edits them in place, works on them unchanged.  No CPU fallback.
"""
import torch
from torch import nn

from . import _native as N


class _CudaBceOps:
    def fwd(self, feat, weight, bias, labels, m, r, t):
        B, C, E = feat.shape[0], weight.shape[0], feat.shape[1]
        dev = feat.device
        logits = torch.empty((B, C), dtype=torch.float32, device=dev)
        gt = torch.empty((B, C), dtype=torch.uint8, device=dev)
        cosine = torch.empty((B, C), dtype=torch.float32, device=dev)
        inv_nf = torch.empty((B,), dtype=torch.float32, device=dev)
        inv_nw = torch.empty((C,), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            N.check(N.lib.pfc_bce_head_fwd(N.ptr(feat), N.ptr(weight), N.ptr(bias), N.ptr(labels), B, C, E, m, r, t,
                                           N.ptr(logits), N.ptr(gt), N.ptr(cosine), N.ptr(inv_nf), N.ptr(inv_nw),
                                           torch.cuda.current_stream(dev).cuda_stream), "pfc_bce_head_fwd")
        return logits, gt.bool(), cosine, inv_nf, inv_nw

    def bwd(self, feat, weight, cosine, inv_nf, inv_nw, dlogits, r, t, need_dfeat, need_dbias):
        B, C, E = feat.shape[0], weight.shape[0], feat.shape[1]
        dev = feat.device
        dfeat = torch.empty_like(feat) if need_dfeat else None
        dweight = torch.empty_like(weight)
        dbias = torch.empty((C,), dtype=torch.float32, device=dev) if need_dbias else None
        with torch.cuda.device(dev):
            N.check(N.lib.pfc_bce_head_bwd(N.ptr(feat), N.ptr(weight), N.ptr(cosine), N.ptr(inv_nf), N.ptr(inv_nw),
                                           N.ptr(dlogits), B, C, E, r, t, N.ptr(dfeat), N.ptr(dweight), N.ptr(dbias),
                                           torch.cuda.current_stream(dev).cuda_stream), "pfc_bce_head_bwd")
        return dfeat, dweight, dbias


class _BceCosineHead(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, weight, bias, labels, m, r, t, ops):
        if ops is None:
            if not (feat.is_cuda and weight.is_cuda):
                raise RuntimeError("fedfr_b200.BCE_module needs CUDA tensors on an sm_100 device (no CPU fallback)")
            ops = _CudaBceOps()
        f = feat.detach().to(torch.float32).contiguous()
        w = weight.detach().to(torch.float32).contiguous()
        b = None if bias is None else bias.detach().to(torch.float32).contiguous()
        y = labels.to(device=f.device, dtype=torch.long).contiguous()
        logits, gt, cosine, inv_nf, inv_nw = ops.fwd(f, w, b, y, float(m), float(r), float(t))
        ctx.save_for_backward(f, w, cosine, inv_nf, inv_nw)
        ctx.ops, ctx.r, ctx.t, ctx.has_bias = ops, float(r), float(t), bias is not None
        ctx.in_dtypes = (feat.dtype, weight.dtype, None if bias is None else bias.dtype)
        ctx.mark_non_differentiable(gt)
        return logits, gt

    @staticmethod
    def backward(ctx, dlogits, _dgt):
        f, w, cosine, inv_nf, inv_nw = ctx.saved_tensors
        need_f, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2]
        dfeat, dweight, dbias = ctx.ops.bwd(f, w, cosine, inv_nf, inv_nw, dlogits.to(torch.float32).contiguous(), ctx.r, ctx.t,
                                            need_f, need_b)
        df = dfeat.to(ctx.in_dtypes[0]) if need_f else None
        dw = dweight.to(ctx.in_dtypes[1]) if need_w else None
        db = dbias.to(ctx.in_dtypes[2]) if need_b else None
        return df, dw, db, None, None, None, None, None


class BCE_module(nn.Module):
    def __init__(self, hidden, n_class, converter_layer=1, m=0.4, r=30.0, t=3, converter=None, _ops=None):
        super(BCE_module, self).__init__()
        if converter is not None:
            self.converter = converter
        elif converter_layer == 1:          # client.py:30-35
            layer = nn.Linear(hidden, hidden)
            nn.init.eye_(layer.weight)
            nn.init.constant_(layer.bias, 0.0)
            self.converter = nn.Sequential(layer)
        else:                               # client.py:36-37 uses backbones.BottleBlock(hidden, 4): pass it as `converter`
            raise NotImplementedError("converter_layer != 1: pass the reference's BottleBlock as converter=")
        self.weight = nn.Parameter(torch.normal(0, 0.01, (n_class, hidden)))
        self.bias = nn.Parameter(torch.zeros(n_class))
        self.g_func = lambda x: (2 * (((x + 1) / 2).pow(t)) - 1)       # kept for callers that read it (client.py:40)
        self.n_class = n_class
        self.hidden = hidden
        self.m = m
        self.r = r
        self.t = t
        self._ops = _ops

    def forward(self, x, labels):
        feat = self.converter(x)
        return _BceCosineHead.apply(feat, self.weight, self.bias, labels, self.m, self.r, self.t, self._ops)

    def initialize(self, fc):
        self.weight.data = fc.clone()
