"""CPU kernel provider for HOST-LOGIC tests only: implements the ``ops`` interface of
``fedfr_b200.PartialFC`` with the oracle so the collective plumbing (gathers, stats exchange, reduce-scatter,
optimizer surgery, update) can run on gloo without a GPU.  Never used by the product."""
import numpy as np
import torch

from oracle import partial_fc_oracle as O


class OracleOps:
    device = "cpu"

    def remap_labels(self, total_label, class_start, num_local):
        return torch.from_numpy(O.remap_labels(total_label.numpy(), class_start, num_local))

    def sample(self, local_label, perm, num_sample):
        idx = O.sample_index(local_label.numpy(), perm.numpy(), num_sample)
        local_label.copy_(torch.from_numpy(O.relabel_to_sample(local_label.numpy(), idx)))
        return torch.from_numpy(idx)

    def gather_rows2(self, weight, weight_mom, index):
        return weight[index].clone(), weight_mom[index].clone()

    def scatter_rows2(self, weight, weight_mom, index, sub_w, sub_m):
        weight_mom[index] = sub_m
        weight[index] = sub_w

    def normalize(self, sub_weight):
        w_hat, n = O.normalize_rows(sub_weight)
        return w_hat, 1.0 / n.squeeze(1)

    def sgd_step(self, weight, weight_mom, grad, index, lr, momentum, dampening, weight_decay, nesterov, prenormalize):
        assert dampening == 0 and not nesterov
        rows = slice(None) if index is None else index
        w, m = O.sgd_momentum_step(weight[rows], weight_mom[rows], grad, lr, momentum, weight_decay)
        weight[rows] = w
        weight_mom[rows] = m
        return None

    def cast_features(self, x):
        return x

    _KIND = {0: "cosface", 1: "arcface"}

    def fwd_stats(self, x, w_hat, label, s, m, margin_kind=0):
        z = O.margin_logits(x, w_hat, label, s, m, self._KIND[margin_kind])
        mx = z.max(dim=1)[0]
        se = torch.exp(z - mx[:, None]).sum(dim=1)
        tz = torch.zeros_like(mx)
        rows = torch.nonzero(label >= 0, as_tuple=True)[0]
        tz[rows] = z[rows, label[rows]]
        return torch.stack([mx, se, tz], dim=1)

    def finalize(self, g):
        M = g[:, :, 0].max(dim=0)[0]
        S = (g[:, :, 1] * torch.exp(g[:, :, 0] - M[None])).sum(dim=0)
        tz = g[:, :, 2].sum(dim=0)
        p = torch.exp(tz - M) / S
        return M, S, -(p.clamp_min(O.PROB_FLOOR).log().mean())

    def bwd(self, x, w_hat, inv_norm, label, row_max, row_sum, s, m, inv_total_batch, dw, accumulate, margin_kind=0):
        z = O.margin_logits(x, w_hat, label, s, m, self._KIND[margin_kind])
        slope = O.target_slope(x, w_hat, label, m, self._KIND[margin_kind])
        out = O.shard_backward(x, z, w_hat, (1.0 / inv_norm)[:, None], label, row_max, row_sum, s, round(1.0 / inv_total_batch), slope)
        if accumulate:
            dw += out.dw
        else:
            dw.copy_(out.dw)
        return out.dx
