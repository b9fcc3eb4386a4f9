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


def margin_params(margin_softmax):
    """(s, m) of a CosFace-like object: this package's class or the reference's ``losses.CosFace``."""
    if type(margin_softmax).__name__ != "CosFace" or not hasattr(margin_softmax, "s") or not hasattr(margin_softmax, "m"):
        raise NotImplementedError(
            f"margin_softmax of type {type(margin_softmax).__name__} is not supported by the fused kernels: "
            "only CosFace(s, m) (losses.py:17-29) is implemented; there is no unfused fallback")
    return float(margin_softmax.s), float(margin_softmax.m)
