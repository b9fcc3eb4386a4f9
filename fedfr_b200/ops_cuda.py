"""Tensor-level wrappers over the C ABI: the only kernel provider of the product path.

Each method takes/returns torch CUDA tensors, passes raw pointers + the current stream to
``libfedfr_b200.so`` and raises ``RuntimeError`` on any non-zero return code.
"""
import os

import torch

from . import _native as N

# backward flavour of the tensor path: "prob" (default) keeps the unnormalised probabilities of the forward GEMM and
# needs no recomputation GEMM; "recompute" re-derives the logits in the backward (no [Bt, Cs] bf16 workspace, no
# restriction on s * |x|, see include/fedfr_b200.h).
BWD_MODE = os.environ.get("FEDFR_BWD_MODE", "prob")


RANGE_LIMIT_NATS = 80.0      # s * |x_i| beyond this leaves the exponent window described in include/fedfr_b200.h
RANGE_SLOTS = 4
_RANGE_FLAGS = {}            # device index -> pinned int32[RANGE_SLOTS, 2] shared by every CudaOps of that device


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


class CudaOps:
    """Kernel provider backed by the sm_100a extension.  ``path`` = PATH_TENSOR (bf16 tcgen05) or
    PATH_CHECK (fp32 SIMT check mode)."""

    def __init__(self, device, path=N.PATH_TENSOR):
        if not torch.cuda.is_available():
            raise RuntimeError("fedfr_b200 needs a CUDA device (sm_100); there is no CPU fallback")
        self.device = torch.device(device)
        self.path = path
        with torch.cuda.device(self.device):
            N.check(N.lib.pfc_query_device(self.device.index or 0, None, None, None), "pfc_query_device")
        self._ws = {}
        self.bwd_mode = BWD_MODE if path == N.PATH_TENSOR else "recompute"
        if self.bwd_mode not in ("prob", "recompute"):
            raise ValueError("FEDFR_BWD_MODE must be 'prob' or 'recompute'")
        self._prob = None       # (workspace, offset, bytes, bt, cs, emb, token) of the last stored-probability forward
        self._sub_rows = None   # row count of the persistent sampled sub-shard buffers
        self.last_token = 0     # bumped by every forward: step buffers (x_hat, w_hat, inv_norm, P) belong to the newest one
        # range guard of the stored-probability path: sticky (s |x| out of the exponent window, row sum of 0) int pairs in
        # pinned host memory that the kernels set every step; RANGE_SLOTS pairs, so that a caller can give every step its
        # own pair and read it back a fixed number of steps later (the same decision on every rank, no synchronisation)
        self._range_flag = _RANGE_FLAGS.get(self.device.index or 0)
        if self._range_flag is None and path == N.PATH_TENSOR:
            self._range_flag = torch.zeros((RANGE_SLOTS, 2), dtype=torch.int32).pin_memory()
            _RANGE_FLAGS[self.device.index or 0] = self._range_flag
        if path == N.PATH_TENSOR:
            self.set_range_slot(0)

    def set_range_slot(self, slot):
        """The stored-probability kernels launched from now on report into pair ``slot``."""
        with torch.cuda.device(self.device):
            N.check(N.lib.pfc_set_range_flag(self._range_flag.data_ptr() + 8 * int(slot), RANGE_LIMIT_NATS), "pfc_set_range_flag")

    def read_range_slot(self, slot):
        """Pair ``slot`` as a bit mask (1: s |x| beyond the window, 2: a row sum of 0), cleared.  Plain read of pinned host
        memory: the caller makes sure the step that wrote it has completed (an event), or accepts a late answer."""
        f = self._range_flag[int(slot)]
        v = int(f[0]) | (int(f[1]) << 1)
        if v:
            f.zero_()
        return v

    # ------------------------------------------------------------------ helpers
    def _persist(self, name, shape, dtype):
        """Per-step scratch tensor with a stable address (the backward is replayed as a CUDA graph keyed on its
        pointer arguments, so step-to-step address churn would force re-captures)."""
        key = (name, tuple(shape), dtype)
        t = self._ws.get(key)
        if t is None:
            t = torch.empty(tuple(shape), dtype=dtype, device=self.device)
            self._ws[key] = t
        return t

    def _buf(self, key, nbytes):
        t = self._ws.get(key)
        if t is None or t.numel() < nbytes:
            t = torch.empty(max(int(nbytes), 1024), dtype=torch.uint8, device=self.device)
            self._ws[key] = t
        return t

    # ------------------------------------------------------------------ integer side
    def remap_labels(self, total_label, class_start, num_local):
        out = torch.empty_like(total_label)
        N.check(N.lib.pfc_remap_labels(N.ptr(total_label), total_label.numel(), int(class_start), int(num_local), N.ptr(out),
                                       _stream(self.device)), "pfc_remap_labels")
        return out

    def remap_labels_(self, total_label, class_start, num_local):
        """In place (the kernel is elementwise): what partial_fc.py:91-93 does to the gathered labels."""
        N.check(N.lib.pfc_remap_labels(N.ptr(total_label), total_label.numel(), int(class_start), int(num_local), N.ptr(total_label),
                                       _stream(self.device)), "pfc_remap_labels")
        return total_label

    def sample(self, local_label, perm, num_sample):
        """In place on ``local_label`` and ``perm``; returns the sorted ``index`` tensor (partial_fc.py:95-104).  The index
        lives in step scratch (valid until the next call), like every other per-step buffer of this provider."""
        num_local = perm.numel()
        cap = max(int(num_sample), min(local_label.numel(), num_local))
        index = self._persist("sample_index", (cap,), torch.int64)
        n_index = self._persist("sample_n_index", (1,), torch.int64)
        ws = self._buf("sample", N.lib.pfc_sample_workspace_bytes(num_local))
        N.check(N.lib.pfc_sample_index(N.ptr(local_label), local_label.numel(), N.ptr(perm), num_local, int(num_sample), N.ptr(index),
                                       N.ptr(n_index), N.ptr(ws), ws.numel(), _stream(self.device)), "pfc_sample_index")
        if num_sample >= local_label.numel():
            return index[:num_sample]          # positives cannot outnumber num_sample: size known without a sync
        return index[: int(n_index.item())]

    def gather_rows2(self, weight, weight_mom, index):
        """(weight[index], weight_mom[index]) into step scratch with stable addresses (a new Parameter wraps the same storage
        every step: the forward / backward graphs keyed on these pointers are captured once)."""
        emb = weight.shape[1]
        n = index.numel()
        if self._sub_rows in (None, n):         # the planned num_sample; a positives-overflow step (other sizes) gets fresh tensors
            self._sub_rows = n
            sub_w = self._persist("sub_w", (n, emb), torch.float32)
            sub_m = self._persist("sub_m", (n, emb), torch.float32)
        else:
            sub_w = torch.empty((n, emb), dtype=torch.float32, device=self.device)
            sub_m = torch.empty_like(sub_w)
        N.check(N.lib.pfc_gather_rows2(N.ptr(weight), N.ptr(weight_mom), N.ptr(index), index.numel(), emb, N.ptr(sub_w), N.ptr(sub_m),
                                       _stream(self.device)), "pfc_gather_rows2")
        return sub_w, sub_m

    def scatter_rows2(self, weight, weight_mom, index, sub_w, sub_m):
        N.check(N.lib.pfc_scatter_rows2(N.ptr(weight), N.ptr(weight_mom), N.ptr(index), index.numel(), weight.shape[1], N.ptr(sub_w),
                                        N.ptr(sub_m), _stream(self.device)), "pfc_scatter_rows2")

    def sgd_step(self, weight, weight_mom, grad, index, lr, momentum, dampening, weight_decay, nesterov, prenormalize):
        """In place torch.optim.SGD step on weight[index] / weight_mom[index] (all rows when index is None).
        Returns (w_hat, inv_norm) of the updated rows when ``prenormalize`` (unsampled, tensor path), else None."""
        n, emb = grad.shape
        w_hat = inv = None
        if prenormalize and index is None and self.path == N.PATH_TENSOR:
            w_hat = self._w_hat_buffer(n, emb)
            inv = self._persist("inv_norm", (n,), torch.float32)
        N.check(N.lib.pfc_sgd_step(N.ptr(weight), N.ptr(weight_mom), N.ptr(grad), N.ptr(index), n, emb, float(lr), float(momentum),
                                   float(dampening), float(weight_decay), 1 if nesterov else 0, N.ptr(w_hat), N.ptr(inv), _stream(self.device)),
                "pfc_sgd_step")
        return (w_hat, inv) if w_hat is not None else None

    # ------------------------------------------------------------------ floating point side
    def normalize(self, sub_weight):
        """-> (w_hat operand, inv_norm).  bf16 on the tensor path, fp32 in check mode."""
        n, emb = sub_weight.shape
        inv = self._persist("inv_norm", (n,), torch.float32)
        if self.path == N.PATH_CHECK:
            w_hat = torch.empty((n, emb), dtype=torch.float32, device=self.device)
            N.check(N.lib.pfc_normalize_rows(N.ptr(sub_weight), None, n, emb, None, N.ptr(w_hat), N.ptr(inv), _stream(self.device)),
                    "pfc_normalize_rows")
        else:
            key = ("w_hat", n, emb)
            w_hat = self._ws.get(key)
            if w_hat is None:
                self._ws = {k: v for k, v in self._ws.items() if not (isinstance(k, tuple) and k[0] == "w_hat")}
                w_hat = torch.empty((n, emb), dtype=torch.bfloat16, device=self.device)
                self._ws[key] = w_hat
            N.check(N.lib.pfc_normalize_rows(N.ptr(sub_weight), None, n, emb, N.ptr(w_hat), None, N.ptr(inv), _stream(self.device)),
                    "pfc_normalize_rows")
        return w_hat, inv

    def _w_hat_buffer(self, n, emb):
        if self.path == N.PATH_CHECK:
            return self._persist("w_hat_f32", (n, emb), torch.float32)
        key = ("w_hat", n, emb)
        w_hat = self._ws.get(key)
        if w_hat is None:
            self._ws = {k: v for k, v in self._ws.items() if not (isinstance(k, tuple) and k[0] == "w_hat")}
            w_hat = torch.empty((n, emb), dtype=torch.bfloat16, device=self.device)
            self._ws[key] = w_hat
        return w_hat

    def normalize_fwd_stats(self, sub_weight, x_hat, label, s, m, margin_kind=0, bwd_mode=None):
        """normalize(sub_weight) fused with fwd_stats (one graph: the normalisation of class chunk k+1 runs under the
        logits kernel of chunk k).  -> (w_hat, inv_norm, stats [Bt, 3])."""
        n, emb = sub_weight.shape
        bt = x_hat.shape[0]
        w_hat = self._w_hat_buffer(n, emb)
        inv = self._persist("inv_norm", (n,), torch.float32)
        n_part = self._const("num_partials", bt, n, emb)
        part = self._persist("part", (2, n_part, bt), torch.float32)
        tz = self._persist("target_logit", (bt,), torch.float32)
        st = _stream(self.device)
        self._prob = None
        if (bwd_mode or self.bwd_mode) == "prob":
            pws, off, nbytes = self._prob_ws(bt, n, emb)
            N.check(N.lib.pfc_normalize_fwd_prob(N.ptr(sub_weight), None, N.ptr(x_hat), N.ptr(label), bt, n, emb, float(s), float(m), int(margin_kind),
                                                 N.ptr(w_hat), N.ptr(inv), N.ptr(part[0]), N.ptr(part[1]), N.ptr(tz), pws.data_ptr() + off, nbytes, st),
                    "pfc_normalize_fwd_prob")
            self._prob = (pws, off, nbytes, bt, n, emb, self.last_token + 1)
        else:
            N.check(N.lib.pfc_normalize_fwd_stats(N.ptr(sub_weight), None, N.ptr(x_hat), N.ptr(label), bt, n, emb, float(s), float(m), int(margin_kind),
                                                  N.ptr(w_hat), N.ptr(inv), N.ptr(part[0]), N.ptr(part[1]), N.ptr(tz), self.path, st),
                    "pfc_normalize_fwd_stats")
        stats = self._persist("stats", (bt, 3), torch.float32)
        N.check(N.lib.pfc_merge_stats(N.ptr(part[0]), N.ptr(part[1]), N.ptr(tz), n_part, bt, N.ptr(stats), st), "pfc_merge_stats")
        self.last_token += 1
        return w_hat, inv, stats

    def _const(self, what, bt, cs, emb):
        """Shape-only quantities of the library (slot counts, workspace sizes): asked once per shape, not once per step."""
        key = ("const", what, bt, cs, emb, self.path)
        v = self._ws.get(key)
        if v is None:
            if what == "num_partials":
                v = N.lib.pfc_fwd_num_partials(bt, cs, emb, self.path)
            elif what == "prob_ws":
                v = N.lib.pfc_prob_workspace_bytes(bt, cs, emb)
            elif what == "bwd_prob_ws":
                v = N.lib.pfc_bwd_prob_workspace_bytes(bt, cs, emb)
            else:
                v = N.lib.pfc_bwd_workspace_bytes(bt, cs, emb, self.path)
            self._ws[key] = v
        return v

    def _prob_ws(self, bt, cs, emb):
        nbytes = self._const("prob_ws", bt, cs, emb)
        ws = self._buf("prob", nbytes + 1024)
        return ws, (-ws.data_ptr()) % 1024, nbytes

    def cast_features(self, total_features):
        if self.path == N.PATH_CHECK:
            return total_features
        x = self._persist("x_hat", total_features.shape, torch.bfloat16)
        N.check(N.lib.pfc_cast_rows_bf16(N.ptr(total_features), total_features.shape[0], total_features.shape[1], N.ptr(x),
                                         _stream(self.device)), "pfc_cast_rows_bf16")
        return x

    def fwd_stats(self, x_hat, w_hat, label, s, m, margin_kind=0, bwd_mode=None):
        """-> stats [Bt, 3] = (row max, sum-exp at that max, target logit) of this shard."""
        bt, emb = x_hat.shape
        cs = w_hat.shape[0]
        n_part = self._const("num_partials", bt, cs, emb)
        part = self._persist("part", (2, n_part, bt), torch.float32)
        tz = self._persist("target_logit", (bt,), torch.float32)
        st = _stream(self.device)
        self._prob = None
        if (bwd_mode or self.bwd_mode) == "prob":     # w == NULL: w_hat / inv_norm are already valid
            pws, off, nbytes = self._prob_ws(bt, cs, emb)
            inv = self._persist("inv_norm", (cs,), torch.float32)
            N.check(N.lib.pfc_normalize_fwd_prob(None, None, N.ptr(x_hat), N.ptr(label), bt, cs, emb, float(s), float(m), int(margin_kind), N.ptr(w_hat),
                                                 N.ptr(inv), N.ptr(part[0]), N.ptr(part[1]), N.ptr(tz), pws.data_ptr() + off, nbytes, st),
                    "pfc_normalize_fwd_prob")
            self._prob = (pws, off, nbytes, bt, cs, emb, self.last_token + 1)
        else:
            N.check(N.lib.pfc_fwd_stats(N.ptr(x_hat), N.ptr(w_hat), N.ptr(label), bt, cs, emb, float(s), float(m), int(margin_kind), N.ptr(part[0]),
                                        N.ptr(part[1]), N.ptr(tz), self.path, st), "pfc_fwd_stats")
        stats = torch.empty((bt, 3), dtype=torch.float32, device=self.device)
        N.check(N.lib.pfc_merge_stats(N.ptr(part[0]), N.ptr(part[1]), N.ptr(tz), n_part, bt, N.ptr(stats), st), "pfc_merge_stats")
        self.last_token += 1
        return stats

    def finalize(self, gathered_stats):
        """[W, Bt, 3] -> (row_max [Bt], row_sum [Bt], loss 0-d)."""
        w, bt, _ = gathered_stats.shape
        row_max = self._persist("row_max", (bt,), torch.float32)
        row_sum = self._persist("row_sum", (bt,), torch.float32)
        loss = torch.empty((), dtype=torch.float32, device=self.device)
        N.check(N.lib.pfc_finalize_stats(N.ptr(gathered_stats), w, bt, N.ptr(row_max), N.ptr(row_sum), N.ptr(loss), _stream(self.device)),
                "pfc_finalize_stats")
        return row_max, row_sum, loss

    def bwd(self, x_hat, w_hat, inv_norm, label, row_max, row_sum, s, m, inv_total_batch, dw, accumulate, margin_kind=0, token=None):
        """Writes/accumulates ``dw`` [Cs, E]; returns this shard's partial ``dx`` [Bt, E] (library-owned scratch,
        overwritten by the next step: callers hand out a copy).  ``token`` = ``last_token`` right after the forward this
        backward belongs to (None: the newest forward); the stored probabilities are only used when they are that forward's,
        otherwise the logits are recomputed from the operands passed in."""
        bt, emb = x_hat.shape
        cs = w_hat.shape[0]
        dx = self._persist("dx", (bt, emb), torch.float32)
        if self._prob is not None and self._prob[3:6] == (bt, cs, emb) and (token is None or token == self._prob[6]):   # its forward kept P
            pws, poff, pbytes = self._prob[:3]
            self._prob = None
            nbytes = self._const("bwd_prob_ws", bt, cs, emb)
            ws = self._buf("bwd_prob", nbytes + 1024)
            off = (-ws.data_ptr()) % 1024
            N.check(N.lib.pfc_bwd_prob(N.ptr(x_hat), N.ptr(w_hat), N.ptr(inv_norm), N.ptr(label), N.ptr(row_sum), bt, cs, emb, float(s), float(m),
                                       int(margin_kind), float(inv_total_batch), N.ptr(dx), N.ptr(dw), 1 if accumulate else 0, pws.data_ptr() + poff,
                                       pbytes, ws.data_ptr() + off, ws.numel() - off, _stream(self.device)), "pfc_bwd_prob")
            return dx
        nbytes = self._const("bwd_ws", bt, cs, emb)
        ws = self._buf("bwd", nbytes + 1024)
        off = (-ws.data_ptr()) % 1024
        N.check(N.lib.pfc_bwd(N.ptr(x_hat), N.ptr(w_hat), N.ptr(inv_norm), N.ptr(label), N.ptr(row_max), N.ptr(row_sum), bt, cs, emb, float(s),
                              float(m), int(margin_kind), float(inv_total_batch), N.ptr(dx), N.ptr(dw), 1 if accumulate else 0, ws.data_ptr() + off,
                              ws.numel() - off, self.path, _stream(self.device)), "pfc_bwd")
        return dx
