"""Pairwise-cosine ROC histogram on the sm_100a kernel -- the B200 side of the reference's ``roc_cuda.py``.

``calc_ROC`` has the argument meaning of the numba kernel of that name (roc_cuda.py:14-28); ``roc_histogram`` is the
whole job that ``roc_cuda.py``'s ``__main__`` spreads over worker processes and batches (roc_cuda.py:30-53, :89-108,
:136-139): here the features stay resident in HBM and the batch loop is ONE launch (``sub_offset`` arithmetic in
``include/fedfr_b200.h``); with a process group the sub rows are split over the ranks by pair count and the int64
histograms are summed with one all-reduce.  ``tpr_at_fpr`` gives the numbers ``plot_ROC`` logs (roc_cuda.py:55-72).
Counts are integer-identical to the reference kernel (fp32 products, fp64 sum, k ascending).  No CPU fallback.
"""
import numpy as np
import torch

from . import _native as N

N_BINS = 2001


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


def _as_cuda(t, dtype, device):
    if isinstance(t, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(t))
    return t.to(device=device, dtype=dtype).contiguous()


def calc_ROC(feature, label, subfeature, sublabel, out, sub_offset=0):
    """``out[2*int((<sub_i, feat_j> + 1) * 1000) + (sublabel_i != label_j)] += 1`` for ``sub_offset + i < j``.

    feature fp32 [n, E], label int32 [n], subfeature fp32 [m, E], sublabel int32 [m], out int64 [2001*2] -- CUDA tensors
    on one device; ``out`` is added to in place and returned (the reference's float64 ``out`` is cast to int64 by its
    caller, roc_cuda.py:52)."""
    if not (feature.is_cuda and subfeature.is_cuda and label.is_cuda and sublabel.is_cuda and out.is_cuda):
        raise RuntimeError("fedfr_b200.calc_ROC needs CUDA tensors (no CPU fallback)")
    if feature.dtype != torch.float32 or subfeature.dtype != torch.float32:
        raise TypeError("features must be float32 (roc_cuda.py:44,47)")
    if label.dtype != torch.int32 or sublabel.dtype != torch.int32:
        raise TypeError("labels must be int32 (roc_cuda.py:45,48)")
    if out.dtype != torch.int64 or out.numel() != 2 * N_BINS or not out.is_contiguous():
        raise TypeError("out must be a contiguous int64 tensor of 2001*2 counters")
    if feature.dim() != 2 or subfeature.dim() != 2 or feature.shape[1] != subfeature.shape[1]:
        raise ValueError("feature [n, E] and subfeature [m, E] must share E")
    if label.numel() != feature.shape[0] or sublabel.numel() != subfeature.shape[0]:
        raise ValueError("one label per feature row")
    feature, subfeature, label, sublabel = (t.contiguous() for t in (feature, subfeature, label, sublabel))
    with torch.cuda.device(feature.device):
        N.check(N.lib.pfc_roc_histogram(N.ptr(feature), N.ptr(label), feature.shape[0], N.ptr(subfeature), N.ptr(sublabel),
                                        subfeature.shape[0], int(sub_offset), feature.shape[1], N.ptr(out),
                                        _stream(feature.device)), "pfc_roc_histogram")
    return out


def shard_rows(target_size, n, world_size):
    """Contiguous sub-row ranges [r0, r1) per rank with near-equal pair counts (row i meets n - 1 - i partners)."""
    target_size, n = int(target_size), int(n)
    pairs = np.maximum(n - 1 - np.arange(target_size, dtype=np.int64), 0)
    cum = np.concatenate([[0], np.cumsum(pairs)])
    cuts = [int(np.searchsorted(cum, cum[-1] * r / world_size, side="left")) for r in range(world_size + 1)]
    cuts[0], cuts[-1] = 0, target_size
    for r in range(1, world_size + 1):
        cuts[r] = max(cuts[r], cuts[r - 1])
    return [(cuts[r], cuts[r + 1]) for r in range(world_size)]


def roc_histogram(feature, label, target_size=None, batch_size=None, device=None, group=None, _calc=None):
    """Histogram of all pairs (i, j), i < target_size, i < j  ->  int64 numpy array [2001, 2] (same-ID, different-ID).

    Equals the sum over the reference's batches (``batch_size`` is accepted for call compatibility and does not change
    the result).  With ``torch.distributed`` initialised (or ``group`` given) each rank takes a row range of the sub
    block and the histograms are all-reduced, so every rank returns the total."""
    import torch.distributed as dist
    n = int(feature.shape[0])
    target_size = n if target_size is None else int(target_size)
    calc = _calc or calc_ROC
    if _calc is None:
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        feature, label = _as_cuda(feature, torch.float32, device), _as_cuda(label, torch.int32, device).reshape(-1)
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    r0, r1 = shard_rows(target_size, n, world)[rank]
    out = torch.zeros(2 * N_BINS, dtype=torch.int64, device=feature.device)
    if r1 > r0:
        calc(feature, label, feature[r0:r1], label[r0:r1], out, sub_offset=r0)
    if world > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out.cpu().numpy().reshape(N_BINS, 2)


def tpr_at_fpr(hist):
    """TPR (percent, two decimals) at FPR = 1e-1 ... 1e-6, as ``plot_ROC`` prints them (roc_cuda.py:55-72)."""
    from scipy.interpolate import interp1d
    cum = np.cumsum(np.asarray(hist, dtype=np.int64).reshape(-1, 2), axis=0)
    tot_same, tot_diff = cum[-1, 0], cum[-1, 1]
    tpr = np.concatenate([[1.0], (tot_same - cum[:, 0]) / tot_same])
    fpr = np.concatenate([[1.0], (tot_diff - cum[:, 1]) / tot_diff])
    order = np.argsort(fpr)
    curve = interp1d(fpr[order], tpr[order])
    return [float("%.2f" % (100 * curve(10.0 ** e))) for e in range(-1, -7, -1)]
