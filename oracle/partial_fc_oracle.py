"""CPU oracle for the PartialFC CosFace hot path  --  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  ``fedfr_b200`` never does.

It restates, as closed-form tensor algebra on the CPU (torch fp32/fp64 + numpy for the integer
work), what the reference computes with autograd and ``torch.distributed``:

* shard geometry ............ reference ``partial_fc.py:34-36``
* label ownership / remap ... reference ``partial_fc.py:91-93``
* negative-centre sampling .. reference ``partial_fc.py:94-106``
* row normalisation ......... reference ``partial_fc.py:127`` (``F.normalize``, eps 1e-12)
* logits .................... reference ``partial_fc.py:108-111``
* CosFace margin ............ reference ``losses.py:23-29``
* distributed softmax-CE .... reference ``partial_fc.py:140-166``
* gradients ................. autograd of ``partial_fc.py:168`` written out by hand
* reduce-scatter + xW ....... reference ``partial_fc.py:171-174``
* write-back ................ reference ``partial_fc.py:113-116``

Pinning: ``tests/test_oracle_golden.py`` checks every function here against vectors produced by
running the *unmodified* reference (``/root/reference/partial_fc.py`` + ``losses.py``) through the
shims in ``tests/golden/make_golden.py`` (gloo, world size 1 and 2).  The reference ships no tests or
golden vectors of its own, so those generated fixtures are the pin.

Tie rule for sampling: the reference takes ``topk(perm, k)`` and sorts the indices.  Which of several
entries *equal* to the k-th value survive is implementation defined in torch (CPU: partial sort, CUDA:
radix select + ordered gather).  The oracle uses the CUDA rule -- strictly-greater entries first, then
equal entries in ascending index order -- and the golden vectors are generated tie-free.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

NORMALIZE_EPS = 1e-12      # F.normalize default, partial_fc.py:127
PROB_FLOOR = 1e-30         # clamp_min_ in partial_fc.py:162


# ----------------------------------------------------------------------------------------------
# integer side
# ----------------------------------------------------------------------------------------------
def shard_geometry(num_classes: int, world_size: int, rank: int) -> Tuple[int, int]:
    """(num_local, class_start) of rank ``rank``  -- partial_fc.py:34-35."""
    base, rem = divmod(num_classes, world_size)
    return base + (1 if rank < rem else 0), base * rank + min(rank, rem)


def num_sample_of(sample_rate: float, num_local: int) -> int:
    """partial_fc.py:36."""
    return int(sample_rate * num_local)


def remap_labels(total_label: np.ndarray, class_start: int, num_local: int) -> np.ndarray:
    """Labels owned by this shard become shard-local ids, everything else -1 (partial_fc.py:91-93)."""
    y = np.asarray(total_label, dtype=np.int64)
    mine = (y >= class_start) & (y < class_start + num_local)
    return np.where(mine, y - class_start, -1).astype(np.int64)


def sample_index(local_label: np.ndarray, perm: np.ndarray, num_sample: int) -> np.ndarray:
    """Sorted shard-local ids of the sampled centres (partial_fc.py:95-102).

    ``perm`` is the fp32 uniform draw of length num_local the reference takes from ``torch.rand``.
    """
    y = np.asarray(local_label, dtype=np.int64)
    positive = np.unique(y[y >= 0])
    if num_sample - positive.size < 0:
        return positive
    score = np.array(perm, dtype=np.float32, copy=True)
    score[positive] = np.float32(2.0)
    if num_sample == 0:
        return np.zeros(0, dtype=np.int64)
    kth = np.partition(score, score.size - num_sample)[score.size - num_sample]
    above = np.flatnonzero(score > kth)
    equal = np.flatnonzero(score == kth)[: num_sample - above.size]
    return np.sort(np.concatenate([above, equal])).astype(np.int64)


def relabel_to_sample(local_label: np.ndarray, index: np.ndarray) -> np.ndarray:
    """searchsorted remap of the owned labels into positions of ``index`` (partial_fc.py:104)."""
    y = np.asarray(local_label, dtype=np.int64).copy()
    own = y >= 0
    y[own] = np.searchsorted(index, y[own])
    return y


# ----------------------------------------------------------------------------------------------
# floating point side (single shard)
# ----------------------------------------------------------------------------------------------
def normalize_rows(w: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(w / max(||w||, eps), max(||w||, eps))  -- F.normalize as used at partial_fc.py:127."""
    n = w.norm(dim=1, keepdim=True).clamp_min(NORMALIZE_EPS)
    return w / n, n


def margin_logits(x: torch.Tensor, w_hat: torch.Tensor, label: torch.Tensor, s: float, m: float, margin: str = "cosface") -> torch.Tensor:
    """CosFace (losses.py:23-29): z = s * (x @ w_hat.T - m * onehot(label)).
    ArcFace (losses.py:38-45): z = s * cos(acos(x @ w_hat.T) + m * onehot(label)); the reference runs acos/cos over the
    whole matrix, which is the identity off the target column (to fp32 rounding), so only targets are touched here.
    Rows with label -1 get no margin."""
    z = x @ w_hat.t()
    rows = torch.nonzero(label >= 0, as_tuple=True)[0]
    if margin == "cosface":
        z[rows, label[rows]] -= m
    elif margin == "arcface":
        z[rows, label[rows]] = torch.cos(torch.acos(z[rows, label[rows]]) + m)
    else:
        raise ValueError(margin)
    return z * s


def target_slope(x: torch.Tensor, w_hat: torch.Tensor, label: torch.Tensor, m: float, margin: str = "cosface") -> torch.Tensor:
    """d(margin(cos)) / d(cos) at the target column, per row (1 where the row has no target here).
    CosFace: 1.  ArcFace: d cos(acos(c) + m) / dc = sin(acos(c) + m) / sqrt(1 - c^2)  (autograd of acos_ / cos_)."""
    f = torch.ones(x.shape[0], dtype=x.dtype)
    if margin == "arcface":
        rows = torch.nonzero(label >= 0, as_tuple=True)[0]
        c = (x[rows] * w_hat[label[rows]]).sum(dim=1)
        f[rows] = torch.sin(torch.acos(c) + m) / torch.sqrt(1.0 - c * c)
    return f


@dataclass
class ShardOut:
    loss_rows: torch.Tensor        # [Bt]  p_{i,y_i} contribution of this shard (0 where not owned)
    dx: torch.Tensor               # [Bt,E] this shard's partial d loss / d total_features
    dw: torch.Tensor               # [Cs,E] d loss / d sub_weight
    row_max: torch.Tensor
    row_sumexp: torch.Tensor


def shard_forward_stats(x, w, label, s, m, margin="cosface"):
    """Local (row max, logits) of one shard; the caller reduces max/sum over shards."""
    w_hat, n = normalize_rows(w)
    z = margin_logits(x, w_hat, label, s, m, margin)
    return z, w_hat, n


def shard_backward(x, z, w_hat, n, label, gmax, gsum, s, total_batch, slope=None) -> ShardOut:
    """Hand-written backward of partial_fc.py:140-168 for one shard given the *global* max / sum-exp."""
    p = torch.exp(z - gmax[:, None]) / gsum[:, None]
    rows = torch.nonzero(label >= 0, as_tuple=True)[0]
    loss_rows = torch.zeros(x.shape[0], dtype=x.dtype)
    loss_rows[rows] = p[rows, label[rows]]
    g = p.clone()
    g[rows, label[rows]] -= 1.0
    g /= total_batch
    g *= s                                        # d z / d cos
    if slope is not None:                         # ArcFace: the target column carries d cos(theta + m) / d cos(theta)
        g[rows, label[rows]] *= slope[rows]
    dx = g @ w_hat                                # [Bt,E]
    dw_hat = g.t() @ x                            # [Cs,E]
    radial = (w_hat * dw_hat).sum(dim=1, keepdim=True)
    dw = (dw_hat - w_hat * radial) / n            # backward of F.normalize
    return ShardOut(loss_rows, dx, dw, z.max(dim=1)[0], torch.exp(z - gmax[:, None]).sum(dim=1))


# ----------------------------------------------------------------------------------------------
# the whole step, all shards simulated in one process
# ----------------------------------------------------------------------------------------------
@dataclass
class StepOut:
    loss: torch.Tensor                         # scalar, identical on every rank
    x_grad: List[torch.Tensor]                 # per rank [B,E]   (already multiplied by world size)
    dw: List[torch.Tensor]                     # per rank [Cs,E]  gradient of sub_weight
    index: List[Optional[np.ndarray]]          # per rank sampled ids (None when sample_rate == 1)
    total_label: List[np.ndarray]              # per rank remapped gathered labels [Bt]


def forward_backward(features: Sequence[torch.Tensor], labels: Sequence[torch.Tensor],
                     weights: Sequence[torch.Tensor], num_classes: int, s: float = 64.0, m: float = 0.4,
                     sample_rate: float = 1.0, perms: Optional[Sequence[np.ndarray]] = None,
                     dtype: torch.dtype = torch.float32, margin: str = "cosface") -> StepOut:
    """One ``PartialFC.forward_backward`` over ``W = len(weights)`` simulated ranks.

    ``features[r]`` [B,E] and ``labels[r]`` [B] are rank r's local batch, ``weights[r]`` its full shard
    ``[num_local_r, E]``.  ``perms[r]`` replaces the ``torch.rand`` draw when ``sample_rate < 1``.
    Collectives (partial_fc.py:122,134,142,147,161,173) are carried out arithmetically.
    """
    W = len(weights)
    B = features[0].shape[0]
    Bt = B * W
    x = torch.cat([f.to(dtype) for f in features], dim=0)                  # all_gather :134
    y_all = torch.cat([l.to(torch.int64) for l in labels], dim=0).numpy()   # all_gather :122

    per_rank = []
    for r in range(W):
        num_local, class_start = shard_geometry(num_classes, W, r)
        assert weights[r].shape[0] == num_local
        y = remap_labels(y_all, class_start, num_local)
        index = None
        w = weights[r].to(dtype)
        if int(sample_rate) != 1:
            index = sample_index(y, perms[r], num_sample_of(sample_rate, num_local))
            y = relabel_to_sample(y, index)
            w = w[torch.from_numpy(index)]
        yt = torch.from_numpy(y)
        z, w_hat, n = shard_forward_stats(x, w, yt, s, m, margin)
        per_rank.append((yt, index, z, w_hat, n, target_slope(x, w_hat, yt, m, margin)))

    gmax = torch.stack([pr[2].max(dim=1)[0] for pr in per_rank]).max(dim=0)[0]        # all_reduce MAX :142
    gsum = sum(torch.exp(pr[2] - gmax[:, None]).sum(dim=1) for pr in per_rank)         # all_reduce SUM :147

    outs = [shard_backward(x, z, w_hat, n, yt, gmax, gsum, s, Bt, slope) for (yt, _, z, w_hat, n, slope) in per_rank]
    p_true = sum(o.loss_rows for o in outs)                                            # all_reduce SUM :161
    loss = -(p_true.clamp_min(PROB_FLOOR).log().mean())                                # :162
    dx_sum = sum(o.dx for o in outs)                                                   # reduce_scatter :173
    x_grad = [dx_sum[r * B:(r + 1) * B] * W for r in range(W)]                         # :174
    return StepOut(loss, x_grad, [o.dw for o in outs], [pr[1] for pr in per_rank],
                   [pr[0].numpy() for pr in per_rank])


def _bf16(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.float32).to(torch.bfloat16).to(torch.float64)


def forward_backward_bf16(features: Sequence[torch.Tensor], labels: Sequence[torch.Tensor], weights: Sequence[torch.Tensor],
                          num_classes: int, s: float = 64.0, m: float = 0.4, sample_rate: float = 1.0,
                          perms: Optional[Sequence[np.ndarray]] = None, margin: str = "cosface",
                          w_hats: Optional[Sequence[torch.Tensor]] = None) -> StepOut:
    """``forward_backward`` with the operand roundings of the bf16 tensor path written out, everything else in fp64
    (so that a kernel on the same rounded operands differs by accumulation order only and can be compared ROW BY ROW at
    ~1e-3 instead of 1e-2 on a whole-tensor norm).  Same reference lines as ``forward_backward``; the roundings are those
    of ``fedfr_b200/csrc/tc_kernels.cu``'s stored-probability path:

        x_hat = bf16(x)                     w_hat = bf16(w / max(|w|, eps))   (or ``w_hats[r]``, the kernel's own operand)
        a_i   = s2 |x_hat_i| (1 + 2^-7) - 58                                   (log2 units; any finite a_i is exact maths)
        P_ij  = bf16(fp32(exp2(s2 cos_ij - a_i)))    S_i = sum_j fp32 P_ij (target with margin, before bf16)
        scale_i = fp32(s / (Bt S_i))       xs_i = bf16(x_hat_i scale_i)
        v_i   = bf16(fp32((p_iy S_i - S_i) slope_i))                           (target element of the scratch)
        dx_i  = scale_i sum_j P'_ij w_hat_j        dW_hat_j = sum_i P'_ij xs_i        dW_j = (dW_hat_j - w_hat_j t_j) fp32(1/n_j)
    """
    W = len(weights)
    B = features[0].shape[0]
    Bt = B * W
    s2 = s * math.log2(math.e)
    x = _bf16(torch.cat([f for f in features], dim=0))
    y_all = torch.cat([l.to(torch.int64) for l in labels], dim=0).numpy()
    a = (torch.tensor(s2, dtype=torch.float32) * x.norm(dim=1).float() * 1.0078125 - 58.0).double()
    per_rank, S = [], torch.zeros(Bt, dtype=torch.float64)
    for r in range(W):
        num_local, class_start = shard_geometry(num_classes, W, r)
        y = remap_labels(y_all, class_start, num_local)
        index = None
        w = weights[r].to(torch.float32)
        if int(sample_rate) != 1:
            index = sample_index(y, perms[r], num_sample_of(sample_rate, num_local))
            y = relabel_to_sample(y, index)
            w = w[torch.from_numpy(index)]
        n = w.norm(dim=1, keepdim=True).clamp_min(NORMALIZE_EPS)
        w_hat = _bf16(w_hats[r]) if w_hats is not None else _bf16(w * (1.0 / n))
        yt = torch.from_numpy(y)
        rows = torch.nonzero(yt >= 0, as_tuple=True)[0]
        cos = x @ w_hat.t()
        c_t = cos[rows, yt[rows]].float()
        if margin == "cosface":
            mc_t, slope = c_t - m, torch.ones_like(c_t)
        else:
            cc = c_t.clamp(-1.0, 1.0)
            mc_t, slope = torch.cos(torch.acos(cc) + m), torch.sin(torch.acos(cc) + m) / torch.sqrt((1.0 - cc * cc).clamp_min(1e-12))
        cos[rows, yt[rows]] = mc_t.double()
        P = torch.exp2(cos * s2 - a[:, None]).float()
        S += P.double().sum(dim=1)
        per_rank.append((yt, index, rows, w_hat, n, P, mc_t, slope))
    scale = (torch.tensor(s / Bt, dtype=torch.float64) / S).float().double()
    xs = _bf16(x * scale[:, None])
    dx_sum = torch.zeros(Bt, x.shape[1], dtype=torch.float64)
    logp = torch.zeros(Bt, dtype=torch.float64)
    dws = []
    for (yt, index, rows, w_hat, n, P, mc_t, slope) in per_rank:
        pS = torch.exp2(mc_t.double() * s2 - a[rows])
        logp[rows] = (mc_t.double() * s2 - a[rows]) * math.log(2.0) - torch.log(S[rows])
        Pq = _bf16(P)
        Pq[rows, yt[rows]] = _bf16(((pS - S[rows]) * slope.double()).float())
        dx_sum += (Pq @ w_hat) * scale[:, None]
        dw_hat = Pq.t() @ xs
        t = (w_hat * dw_hat).sum(dim=1, keepdim=True)
        dws.append((dw_hat - w_hat * t) * (1.0 / n).float().double())
    loss = torch.minimum(-logp, torch.full_like(logp, -math.log(PROB_FLOOR))).mean()
    x_grad = [dx_sum[r * B:(r + 1) * B] * W for r in range(W)]
    return StepOut(loss, x_grad, dws, [pr[1] for pr in per_rank], [pr[0].numpy() for pr in per_rank])


def dense_twin_grads(x, w, label, s, m):
    """Second, independent statement of the same maths through autograd: client.py:69-74 (FC_module with
    pre-normalised features) + losses.py:23-29 + F.cross_entropy (client.py:433).  Returns (loss, dx, dw)."""
    x = x.clone().requires_grad_(True)
    w = w.clone().requires_grad_(True)
    cos = x @ torch.nn.functional.normalize(w).t()
    onehot = torch.nn.functional.one_hot(label, w.shape[0]).to(cos.dtype)
    loss = torch.nn.functional.cross_entropy(s * (cos - m * onehot), label)
    loss.backward()
    return loss.detach(), x.grad, w.grad


def sgd_momentum_step(w, mom, grad, lr, momentum=0.9, weight_decay=5e-4):
    """torch.optim.SGD (dampening 0, no nesterov) as driven by partial_fc.py:124-126 -- used to check the
    optimizer surgery and ``update()`` write-back (partial_fc.py:113-116)."""
    g = grad + weight_decay * w
    mom = momentum * mom + g
    return w - lr * mom, mom


# ----------------------------------------------------------------------------------------------
# FedAvg  (server.py:25-46)
# ----------------------------------------------------------------------------------------------
def fedavg_weights(weights: Sequence[float]) -> List[float]:
    """server.py:27 -- Python-float normalisation, later rounded to fp32 when it meets a tensor."""
    tot = sum(weights)
    return [w / tot for w in weights]


def fedpavg(models: Sequence[dict], weights: Sequence[float]) -> dict:
    """server.py:25-34 with numpy: per key, client order, fp32 mul then fp32 add (no FMA).  Integer
    buffers go through float32 exactly like ``python_float * int64_tensor`` does in torch."""
    wn = [np.float32(w) for w in fedavg_weights(weights)]
    out = {}
    for key in models[0]:
        acc = None
        for wi, sd in zip(wn, models):
            v = np.asarray(sd[key])
            term = (v.astype(np.float32) * wi).astype(np.float32)
            acc = term if acc is None else (acc + term).astype(np.float32)
        out[key] = acc
    return out


def fedavg_on_fc(pretrain_fc: np.ndarray, models: Sequence[np.ndarray], weights: Sequence[float], p: float) -> np.ndarray:
    """server.py:36-46."""
    wn = [np.float32(w) for w in fedavg_weights(weights)]
    acc = (np.asarray(models[0], dtype=np.float32) * wn[0]).astype(np.float32)
    for wi, mdl in zip(wn[1:], models[1:]):
        acc = (acc + (np.asarray(mdl, dtype=np.float32) * wi).astype(np.float32)).astype(np.float32)
    if p == 1:
        return acc
    return ((np.float32(1 - p) * np.asarray(pretrain_fc, dtype=np.float32)).astype(np.float32)
            + (np.float32(p) * acc).astype(np.float32)).astype(np.float32)


def spreadout_loss(fc: torch.Tensor, margin: float, mode: str = "sum") -> torch.Tensor:
    """server.py:55-62 (SpreadOut_Module.forward) as written, minus the ``.cuda()`` on the mask: normalise, full
    similarity, off-diagonal entries, relu(sim - margin)^2, sum or mean.  Differentiable (use autograd for the gradient)."""
    w = torch.nn.functional.normalize(fc)
    similarity = torch.matmul(w, w.t())
    off = similarity.masked_select(~torch.eye(len(w), dtype=torch.bool))
    loss = torch.nn.functional.relu(off - margin) ** 2
    return loss.sum() if mode == "sum" else loss.mean()
