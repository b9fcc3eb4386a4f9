"""Self-check of one ``PartialFC.forward_backward`` step at ANY size and world size, on the device.

A chunked restatement of partial_fc.py:127-174 + losses.py:23-45 in plain torch ops (fp32 GEMMs, fp64 sums) that runs
where the tensors live, so that ``bench.py`` can assert parity after its timed loops at the full BASELINE shapes and on
every rank of a multi-GPU job (the CPU oracle under ``oracle/`` finishes only small cases and never travels with the
product).  Two flavours:

* ``emulate=False`` -- the reference's arithmetic on the fp32 operands (what ``oracle/partial_fc_oracle.py`` restates);
  the bf16 tensor path is held to BASELINE's 1e-2 against it.
* ``emulate=True``  -- the same algebra with the operand roundings of the tensor path made explicit (bf16 ``x_hat`` /
  ``w_hat``, bf16 stored probabilities ``P = exp2(s2 cos - a_i)``, bf16 ``x_hat * row_scale``, fp32 target fix-up), so
  that only accumulation order and ``ex2.approx`` differ: rows are compared one by one at 2e-3, which a systematic
  error in the non-target part of ``dw`` / ``dx`` cannot hide behind the large target rows.

Nothing here is on the product path: it is called by ``bench.py`` (after timing) and by ``tests/``.
"""
import math

import torch

LOG2E = 1.4426950408889634
LN2 = 0.6931471805599453
HEADROOM = 58.0           # log2 units, csrc/tc_kernels.cu kProbHeadroom


def _margin(c, m, kind):
    """(margin cosine, slope) of the target column: CosFace losses.py:23-29, ArcFace losses.py:38-45."""
    if kind == 0:
        return c - m, torch.ones_like(c)
    c = c.clamp(-1.0, 1.0)
    th = torch.acos(c)
    return torch.cos(th + m), torch.sin(th + m) * torch.rsqrt((1.0 - c * c).clamp_min(1e-12))


def _all_reduce(t, op, group):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=op, group=group)
    return t


def reference_step(x_total, y_local, w_sub, s, m, kind=0, w_hat=None, emulate=False, dw_rows=None, chunk=32768, group=None):
    """One rank's share of the step.

    x_total [Bt, E] fp32 gathered features; y_local [Bt] int64 column id in ``w_sub`` or -1; w_sub [Cs, E] fp32 rows of
    this rank (sampled rows when sample_rate < 1); ``w_hat``: the bf16 operand the kernels produced (``emulate`` only;
    default = normalize(w_sub) rounded here).  ``dw_rows``: int64 ids of the ``w_sub`` rows whose gradient is wanted
    (None = all, small cases only).  Collectives (all-reduce of row statistics, target logits and dx) run on ``group``
    when torch.distributed is initialised.  Returns a dict: loss (0-d fp64), dx_total [Bt, E] fp64 (sum over ranks of
    s G w_hat; x_grad of rank r = its rows * world_size, partial_fc.py:171-174), dw [len(dw_rows), E] fp64, dw_rows."""
    import torch.distributed as dist
    dev = x_total.device
    Bt, E = x_total.shape
    Cs = w_sub.shape[0]
    n = w_sub.float().norm(dim=1, keepdim=True).clamp_min(1e-12)                 # normalize(), partial_fc.py:127
    if emulate:
        xb = x_total.to(torch.bfloat16).float()
        wb = w_hat.float() if w_hat is not None else (w_sub.float() * (1.0 / n)).to(torch.bfloat16).float()
    else:
        xb = x_total.float()
        wb = w_sub.float() / n
    s2 = s * LOG2E
    own = y_local >= 0
    yl = y_local.clamp_min(0)
    # per-row reference point a_i (log2 units): the tensor path's bound; any finite value gives the same mathematics
    a = s2 * xb.norm(dim=1) * 1.0078125 - HEADROOM if emulate else torch.full((Bt,), 0.0, device=dev)
    c_t = (xb * wb[yl]).sum(dim=1)                                                # plain target cosine (owner rows)
    mc_t, slope = _margin(c_t, m, kind)
    if not emulate:                                                              # exact path: reference point = global row max
        mx = torch.full((Bt,), -float("inf"), device=dev)
        for c0 in range(0, Cs, chunk):
            z = xb @ wb[c0:c0 + chunk].t()
            hit = own & (yl >= c0) & (yl < c0 + chunk)
            z[hit, yl[hit] - c0] = mc_t[hit]
            mx = torch.maximum(mx, z.max(dim=1)[0])
        _all_reduce(mx, dist.ReduceOp.MAX if dist.is_available() else None, group)
        a = mx * s2
    S = torch.zeros(Bt, dtype=torch.float64, device=dev)
    dx = torch.zeros(Bt, E, dtype=torch.float64, device=dev)
    for c0 in range(0, Cs, chunk):
        z = xb @ wb[c0:c0 + chunk].t()
        hit = own & (yl >= c0) & (yl < c0 + chunk)
        z[hit, yl[hit] - c0] = mc_t[hit]
        P = torch.exp2(z * s2 - a[:, None])
        S += P.double().sum(dim=1)
        P[hit, yl[hit] - c0] = 0.0                                               # target handled in fp32 below
        if emulate:
            P = P.to(torch.bfloat16).float()
        dx += (P @ wb[c0:c0 + chunk]).double()
    _all_reduce(S, dist.ReduceOp.SUM, group)
    pS_t = torch.exp2(mc_t * s2 - a).double()                                     # p_iy * S_i  (owner rows)
    logp = torch.where(own, (mc_t.double() * s2 - a.double()) * LN2 - torch.log(S), torch.zeros_like(S))
    _all_reduce(logp, dist.ReduceOp.SUM, group)
    loss = torch.minimum(-logp, torch.full_like(logp, -math.log(1e-30))).mean()   # clamp_min(1e-30), partial_fc.py:162
    g_scale = s / Bt
    scale = g_scale / S                                                          # row scale s / (S_i Bt)
    v_t = (pS_t - S) * slope.double()                                            # target element (p_iy - 1) slope S_i
    if emulate:
        scale = scale.float().double()
        v_t = v_t.float().to(torch.bfloat16).double()
    v_t = torch.where(own, v_t, torch.zeros_like(v_t))
    dx += v_t[:, None] * wb[yl].double()
    dx *= scale[:, None]
    _all_reduce(dx, dist.ReduceOp.SUM, group)
    # dw on the requested rows
    if dw_rows is None:
        dw_rows = torch.arange(Cs, device=dev)
    xs = xb.double() * scale[:, None]
    if emulate:
        xs = xs.float().to(torch.bfloat16).double()
    dw = torch.empty(dw_rows.numel(), E, dtype=torch.float64, device=dev)
    for r0 in range(0, dw_rows.numel(), chunk):
        rows = dw_rows[r0:r0 + chunk]
        z = xb @ wb[rows].t()
        hitm = own[:, None] & (yl[:, None] == rows[None, :])
        P = torch.exp2(z * s2 - a[:, None])
        if emulate:
            P = P.to(torch.bfloat16).float()
        P = torch.where(hitm, v_t[:, None].float().expand_as(P), P).double()
        dwh = P.t() @ xs
        wr = wb[rows].double()
        inv_n = (1.0 / n[rows]).double() if not emulate else (1.0 / n[rows]).float().double()
        dw[r0:r0 + chunk] = (dwh - wr * (wr * dwh).sum(dim=1, keepdim=True)) * inv_n
    return {"loss": loss, "dx_total": dx, "dw": dw, "dw_rows": dw_rows, "w_hat_ref": wb}


def _rows_err(got, ref, rms=False):
    """Per-row relative error |got - ref| / |ref| (floor: 1 % of the rms row norm, so vanishing rows do not divide by ~0):
    its maximum over the rows, or (``rms``) its root mean square."""
    got, ref = got.double(), ref.double()
    if not ref.numel():
        return 0.0
    rn = ref.norm(dim=1)
    floor = 0.01 * float(rn.pow(2).mean().sqrt()) + 1e-30
    e = (got - ref).norm(dim=1) / rn.clamp_min(floor)
    return float(e.pow(2).mean().sqrt()) if rms else float(e.max())


def _rel(got, ref):
    got, ref = got.double(), ref.double()
    return float((got - ref).norm() / (ref.norm() + 1e-7 * ref.numel() ** 0.5))


def check_head_step(head, label, features, x_grad, loss, group=None, dw_stride=1009, tol=1e-2, tol_rows=1e-2, tol_rows_rms=1e-3):
    """Compare the outputs of ONE ``head.forward_backward(label, features, opt)`` (``x_grad``, ``loss`` and the freshly
    written ``head.sub_weight.grad``; the weights must not have been stepped since) with the restatement, on every rank.
    Returns a dict of errors with ``ok``; identical ``loss`` bits on all ranks are part of the check.

    Row tolerances of the bf16-emulating comparison (kernel and restatement round the same values to bf16, but reach them
    by different summation orders, so a value near a rounding boundary may go to the other neighbour -- one ulp = 2^-8):
    * EVERY row of ``dx`` and of the checked ``dw`` rows within ``tol_rows`` = 1e-2 (2.5 ulp): a row dominated by one or two
      rounded elements (target rows; every ``dx`` row early in training; a class that one sample hits hard) moves by their ulps;
    * the ROOT MEAN SQUARE of the per-row errors within ``tol_rows_rms`` = 1e-3, separately for ``dx`` rows, target rows
      and non-target rows of ``dw``: flips are rare and unsigned, so a systematic error of a fraction of a percent in any
      of the three groups (which a whole-tensor norm hides behind the large target rows) fails this."""
    import torch.distributed as dist
    W, rank = head.world_size, head.rank
    dev = head.device
    feats = features.detach().to(dev, torch.float32).contiguous()
    lab = label.detach().to(dev, torch.int64).contiguous()
    if W > 1:
        xt = torch.empty(W * feats.shape[0], feats.shape[1], device=dev)
        yt = torch.empty(W * lab.shape[0], dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(xt, feats, group=group)
        dist.all_gather_into_tensor(yt, lab, group=group)
    else:
        xt, yt = feats, lab
    sampled = int(head.sample_rate) != 1
    y_local = yt - head.class_start                                                # partial_fc.py:91-93
    y_local = torch.where((y_local >= 0) & (y_local < head.num_local), y_local, torch.full_like(y_local, -1))
    if sampled:                                                                    # partial_fc.py:104
        idx = head.index
        pos = torch.searchsorted(idx, y_local.clamp_min(0))
        y_local = torch.where(y_local >= 0, pos, y_local)
        w_sub = head.weight[idx]
    else:
        w_sub = head.weight
    Cs = w_sub.shape[0]
    own_targets = torch.unique(y_local[y_local >= 0])
    rows = torch.unique(torch.cat([torch.arange(0, Cs, dw_stride, device=dev), own_targets]))
    is_target = torch.isin(rows, own_targets)
    dw_got = head.sub_weight.grad[rows]
    B = feats.shape[0]
    out = {"rank": rank, "world": W, "rows_checked": int(rows.numel())}
    w_hat = head._norm[0] if getattr(head, "_norm", None) is not None and head._norm[0].dtype == torch.bfloat16 else None
    for name, emulate in (("fp32", False), ("bf16", True)):
        if emulate and w_hat is None:
            continue
        ref = reference_step(xt, y_local, w_sub, head._s, head._m, head._margin_kind, w_hat=w_hat, emulate=emulate, dw_rows=rows, group=group)
        xg_ref = ref["dx_total"][rank * B:(rank + 1) * B] * W
        out[name] = {
            "loss_rel": abs(float(loss) - float(ref["loss"])) / max(abs(float(ref["loss"])), 1e-30),
            "dx_rel": _rel(x_grad, xg_ref), "dw_rel": _rel(dw_got, ref["dw"]),
            "dx_rows_max": _rows_err(x_grad, xg_ref), "dx_rows_rms": _rows_err(x_grad, xg_ref, rms=True),
            "dw_target_rows_max": _rows_err(dw_got[is_target], ref["dw"][is_target]),
            "dw_target_rows_rms": _rows_err(dw_got[is_target], ref["dw"][is_target], rms=True),
            "dw_other_rows_max": _rows_err(dw_got[~is_target], ref["dw"][~is_target]),
            "dw_other_rows_rms": _rows_err(dw_got[~is_target], ref["dw"][~is_target], rms=True),
        }
    if w_hat is not None:                                                          # normalize(): within one bf16 rounding of the fp32 value
        wn = torch.nn.functional.normalize(w_sub[rows].float())
        out["w_hat_max_rel"] = float(((w_hat[rows].float() - wn).abs() / wn.abs().clamp_min(1e-6)).max())
    lv = loss.detach().reshape(1).to(torch.float32).clone()
    same = True
    if W > 1:
        allv = torch.empty(W, device=dev)
        dist.all_gather_into_tensor(allv, lv, group=group)
        same = bool((allv.view(torch.int32) == allv.view(torch.int32)[0]).all())
    out["loss_identical_on_all_ranks"] = same
    ok = same and out["fp32"]["loss_rel"] < tol and out["fp32"]["dx_rel"] < tol and out["fp32"]["dw_rel"] < tol
    if "bf16" in out:
        b = out["bf16"]
        ok = ok and b["loss_rel"] < 1e-4 and max(b["dx_rows_max"], b["dw_target_rows_max"], b["dw_other_rows_max"]) < tol_rows
        ok = ok and max(b["dx_rows_rms"], b["dw_target_rows_rms"], b["dw_other_rows_rms"]) < tol_rows_rms
        ok = ok and out["w_hat_max_rel"] <= 2.0 ** -8 * 1.01
    out["ok"] = bool(ok)
    if W > 1:                                                                      # every rank must pass
        flag = torch.tensor([1.0 if ok else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        out["ok_all_ranks"] = bool(flag.item() > 0)
    else:
        out["ok_all_ranks"] = out["ok"]
    return out
