"""Dense (single-GPU) margin-softmax head on the fused kernels, with autograd -- the call site FedFR's live client uses.

``client.py`` trains with ``logits = FC_module(x)`` (``matmul(normalize(x), normalize(fc).t())``, client.py:69-74),
``logits = CosFace(s, m)(logits, labels)`` (losses.py:23-29, applied at client.py:377,388,421,430,544) and
``loss = F.cross_entropy(logits, labels); loss.backward()`` (client.py:433-435).  That is exactly
``PartialFC.forward_backward`` at world_size 1 applied to ``normalize(x)``, so the same
kernels serve it; this module only adds the autograd glue so that the client's ``loss.backward()`` flow works unchanged:

    loss = margin_cross_entropy(embeddings, fc.weight, labels, CosFace(s=30, m=0.4))     # replaces the three lines above
    loss.backward()                                                                      # embeddings.grad, fc.weight.grad

The ``[B, C]`` logits are never materialised.  There is no unfused fallback.
"""
import torch

from .losses import margin_params

_OPS = {}
_PROB_MAX_S = 80.0      # ops_cuda.RANGE_LIMIT_NATS: beyond it the head uses the recomputing backward (no domain limit)


def _ops_for(device, check_mode):
    from . import _native as N
    from .ops_cuda import CudaOps
    key = (device.index if device.index is not None else torch.cuda.current_device(), bool(check_mode))
    ops = _OPS.get(key)
    if ops is None:
        ops = CudaOps(torch.device("cuda", key[0]), N.PATH_CHECK if check_mode else N.PATH_TENSOR)
        _OPS[key] = ops
    return ops


class _MarginCrossEntropy(torch.autograd.Function):
    """loss = mean_i CE(s * margin(cos(x_i, w_j)), y_i); backward through normalize(x) and normalize(w)."""

    @staticmethod
    def forward(ctx, x, weight, label, s, m, kind, ops):
        x32 = x.detach().to(torch.float32)
        norm = x32.norm(dim=1, keepdim=True).clamp_min(1e-12)            # F.normalize(x), client.py:70
        x_unit = x32 / norm
        w = weight.detach().to(torch.float32).contiguous()             # the kernels read raw fp32 [C, E] rows (AMP casts, sliced fc views)
        label = label.to(torch.int64).contiguous()
        x_hat = ops.cast_features(x_unit.contiguous())
        # the features are unit vectors here, so the stored-probability window (s |x| <= 80 nats) is a condition on s alone
        mode = "recompute" if s > _PROB_MAX_S else None
        if hasattr(ops, "normalize_fwd_stats"):
            w_hat, inv_norm, stats = ops.normalize_fwd_stats(w, x_hat, label, s, m, kind, **({"bwd_mode": mode} if mode else {}))
        else:
            w_hat, inv_norm = ops.normalize(w)
            stats = ops.fwd_stats(x_hat, w_hat, label, s, m, kind)
        row_max, row_sum, loss = ops.finalize(stats.unsqueeze(0))
        ctx.ops, ctx.margin = ops, (s, m, kind)
        ctx.token = getattr(ops, "last_token", None)
        # row_max / row_sum live in step scratch of `ops`: keep private copies for a backward that runs later.  x_hat, w_hat
        # and inv_norm are step scratch too; backward() rebuilds them if another forward used `ops` in between.
        ctx.save_for_backward(x_unit, norm, x_hat, w_hat, inv_norm, label, row_max.clone(), row_sum.clone(), w)
        ctx.w_shape, ctx.w_dtype, ctx.x_dtype = tuple(w.shape), weight.dtype, x.dtype
        return loss.clone()

    @staticmethod
    def backward(ctx, grad_out):
        x_unit, norm, x_hat, w_hat, inv_norm, label, row_max, row_sum, w = ctx.saved_tensors
        s, m, kind = ctx.margin
        ops = ctx.ops
        dw = torch.empty(ctx.w_shape, dtype=torch.float32, device=x_unit.device)
        if ctx.token is not None and ops.last_token != ctx.token:
            # another forward (a second head of the same shape, a PartialFC on this device) has reused the step buffers:
            # rebuild this head's operands; the stored probabilities are gone, so the backward recomputes the logits
            x_hat = ops.cast_features(x_unit.contiguous())
            w_hat, inv_norm = ops.normalize(w)
            ops.last_token += 1                              # ... which in turn invalidates whoever owned the buffers until now
            d_unit = ops.bwd(x_hat, w_hat, inv_norm, label, row_max, row_sum, s, m, 1.0 / x_unit.shape[0], dw, False, kind, token=-1)
        elif ctx.token is not None:
            d_unit = ops.bwd(x_hat, w_hat, inv_norm, label, row_max, row_sum, s, m, 1.0 / x_unit.shape[0], dw, False, kind, token=ctx.token)
        else:
            d_unit = ops.bwd(x_hat, w_hat, inv_norm, label, row_max, row_sum, s, m, 1.0 / x_unit.shape[0], dw, False, kind)
        d_unit = d_unit.clone()                              # dx is step scratch as well
        # backward of F.normalize(x):  dx = (d - x_unit (x_unit . d)) / |x|
        dx = (d_unit - x_unit * (x_unit * d_unit).sum(dim=1, keepdim=True)) / norm
        g = grad_out.to(torch.float32)
        return (dx * g).to(ctx.x_dtype), (dw * g).to(ctx.w_dtype), None, None, None, None, None


def margin_cross_entropy(x, weight, label, margin_softmax, check_mode=False, _ops=None):
    """``F.cross_entropy(margin_softmax(F.linear(F.normalize(x), F.normalize(weight)), label), label)`` without the
    logits.  ``x`` [B, E], ``weight`` [C, E] fp32 CUDA tensors (either may require grad), ``label`` int64 [B] in [0, C),
    ``margin_softmax`` a ``CosFace(s, m)`` / ``ArcFace(s, m)`` descriptor (this package's or the reference's).
    One head at a time per device and stream: the step scratch of the kernel provider is shared per device (a second forward
    in between makes this head's backward rebuild its operands, see ``_MarginCrossEntropy.backward``); heads running
    concurrently on DIFFERENT CUDA streams of one device are not supported."""
    s, m, kind = margin_params(margin_softmax)
    if x.dim() != 2 or weight.dim() != 2 or x.shape[1] != weight.shape[1] or label.shape[0] != x.shape[0]:
        raise ValueError("margin_cross_entropy: x [B, E], weight [C, E], label [B] expected")
    if weight.device != x.device or label.device != x.device:
        raise ValueError("margin_cross_entropy: x, weight and label must live on the same device")
    if not (weight.is_floating_point() and x.is_floating_point()):
        raise TypeError("margin_cross_entropy: x and weight must be floating point tensors")
    if not (s > 0 and s == s and s != float("inf")):
        raise ValueError("margin_cross_entropy: the scale s must be positive and finite")
    ops = _ops if _ops is not None else _ops_for(x.device, check_mode)
    return _MarginCrossEntropy.apply(x, weight, label, s, m, kind, ops)


class MarginSoftmaxHead(torch.nn.Module):
    """``FC_module`` (client.py:63-83) + margin + cross-entropy in one module: holds ``fc`` [n_class, embedding_size]
    (same parameter name and init as client.py:66-67) and returns the loss."""

    def __init__(self, n_class, margin_softmax, embedding_size=512, check_mode=False):
        super().__init__()
        self.fc = torch.nn.Parameter(torch.normal(0, 0.01, (n_class, embedding_size)))
        self.margin_softmax = margin_softmax
        self.check_mode = check_mode

    def forward(self, x, label):
        return margin_cross_entropy(x, self.fc, label, self.margin_softmax, self.check_mode)
