"""Hard-negative mining by similarity threshold on the sm_100a kernel (``csrc/hardneg.cu``).

``similar_columns(a, b, threshold)`` is the array part of ``Client.choose_hard_negative_2`` (client.py:208-215):
``similarity = matmul(local_feats, pretrained_feats.t())`` followed by the union over rows of
``torch.where(similarity > threshold)[1]`` -- returned as the sorted int64 index array the reference builds with
``sorted(reduce(np.union1d, ...))``.  ``hard_negative_ids`` is the FC-based variant (client.py:232-235), which
normalises both operands first.  The [n_a, n_b] similarity matrix (n_b = 420 k public images in FedFR) is never
materialised, so the reference's 100-slice loop "to prevent out of mem of RAM" has no counterpart.  No CPU fallback.
"""
import numpy as np
import torch

from . import _native as N


def _hit_mask(a, b, threshold):
    if not (a.is_cuda and b.is_cuda):
        raise RuntimeError("fedfr_b200.similar_columns needs CUDA tensors on an sm_100 device (no CPU fallback)")
    if a.dim() != 2 or b.dim() != 2 or a.shape[1] != b.shape[1]:
        raise ValueError("a [n_a, E] and b [n_b, E] must share E")
    a = a.detach().to(torch.float32).contiguous()
    b = b.detach().to(device=a.device, dtype=torch.float32).contiguous()
    hit = torch.empty((b.shape[0],), dtype=torch.uint8, device=a.device)
    with torch.cuda.device(a.device):
        N.check(N.lib.pfc_similar_columns(N.ptr(a), a.shape[0], N.ptr(b), b.shape[0], a.shape[1], float(threshold), N.ptr(hit),
                                          torch.cuda.current_stream(a.device).cuda_stream), "pfc_similar_columns")
    return hit


def similar_columns(a, b, threshold=0.2, _mask_fn=None):
    """Sorted unique column ids j with ``max_i <a_i, b_j> > threshold`` (numpy int64), client.py:208-215."""
    hit = (_mask_fn or _hit_mask)(a, b, threshold)
    return torch.nonzero(hit, as_tuple=True)[0].cpu().numpy().astype(np.int64)


def hard_negative_ids(self_fc, pretrain_fc, threshold=0.2, _mask_fn=None):
    """``torch.unique(torch.where(normalize(self_fc) @ normalize(pretrain_fc).t() > threshold)[1]).numpy()``, client.py:232-235."""
    if not isinstance(threshold, float):
        raise NotImplementedError("integer thresholds (top-k) raise in the reference as well (client.py:236-238)")
    return similar_columns(torch.nn.functional.normalize(self_fc), torch.nn.functional.normalize(pretrain_fc), threshold, _mask_fn)
