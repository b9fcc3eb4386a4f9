"""Margin descriptors with the constructor surface of the reference's ``losses.py``.

``CosFace(s, m)`` (losses.py:17-29) is consumed by :class:`fedfr_b200.PartialFC` as a *descriptor*: the
margin and scale are fused into the epilogue of the logits kernels, the dense ``[Bt, Cs]`` logits tensor the
reference edits in place never exists.  Calling the object on materialised logits (the dense twin of
client.py:430) runs the same arithmetic through a CUDA elementwise kernel.
"""
import torch


class CosFace(torch.nn.Module):
    def __init__(self, s=64.0, m=0.40):
        super().__init__()
        self.s = s
        self.m = m

    def forward(self, cosine, label):
        """losses.py:23-29 -- in place ``cosine[i, label[i]] -= m`` for ``label[i] != -1``; returns ``cosine * s``."""
        from . import _native as N
        if not cosine.is_cuda:
            raise RuntimeError("fedfr_b200.losses.CosFace runs on CUDA tensors only (no CPU fallback)")
        assert cosine.dtype == torch.float32 and cosine.is_contiguous() and label.dtype == torch.int64
        out = torch.empty_like(cosine)
        st = torch.cuda.current_stream(cosine.device).cuda_stream
        N.check(N.lib.pfc_cosface_dense(N.ptr(cosine), N.ptr(label.contiguous()), cosine.shape[0], cosine.shape[1], float(self.s),
                                        float(self.m), N.ptr(out), st), "pfc_cosface_dense")
        return out


class ArcFace(torch.nn.Module):
    """losses.py:32-45: ``s * cos(acos(cosine) + m)`` on the target column.  ``PartialFC`` / ``margin_cross_entropy`` read
    the object as a descriptor and fuse it into the logits epilogues (target cosine clamped to [-1, 1] first); calling it
    on materialised logits (client.py:430 with ``--loss ArcFace``) runs the reference's arithmetic in one elementwise
    kernel, in place on ``cosine`` and returning it, exactly like the reference."""

    def __init__(self, s=64.0, m=0.5):
        super().__init__()
        self.s = s
        self.m = m

    def forward(self, cosine, label):
        from . import _native as N
        if not cosine.is_cuda:
            raise RuntimeError("fedfr_b200.losses.ArcFace runs on CUDA tensors only (no CPU fallback)")
        assert cosine.dtype == torch.float32 and cosine.is_contiguous() and label.dtype == torch.int64
        st = torch.cuda.current_stream(cosine.device).cuda_stream
        N.check(N.lib.pfc_arcface_dense(N.ptr(cosine), N.ptr(label.contiguous()), cosine.shape[0], cosine.shape[1], float(self.s),
                                        float(self.m), st), "pfc_arcface_dense")
        return cosine


_MARGIN_KINDS = {"CosFace": 0, "ArcFace": 1}      # PFC_MARGIN_COSFACE / PFC_MARGIN_ARCFACE


def margin_params(margin_softmax):
    """(s, m, margin_kind) of a margin object: this package's classes or the reference's ``losses.CosFace`` /
    ``losses.ArcFace`` (matched by class name, read as descriptors, never called)."""
    name = type(margin_softmax).__name__
    if name not in _MARGIN_KINDS or not hasattr(margin_softmax, "s") or not hasattr(margin_softmax, "m"):
        raise NotImplementedError(
            f"margin_softmax of type {name} is not supported by the fused kernels: CosFace(s, m) (losses.py:17-29) and "
            "ArcFace(s, m) (losses.py:32-45) are implemented; there is no unfused fallback")
    return float(margin_softmax.s), float(margin_softmax.m), _MARGIN_KINDS[name]
