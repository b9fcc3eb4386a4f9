"""``PartialFC`` -- class-sharded, optionally sampled CosFace margin-softmax head on B200.

Drop-in for the reference's ``partial_fc.PartialFC`` (partial_fc.py:11-176): same constructor
arguments, attributes (``weight``, ``weight_mom``, ``sub_weight``, ``sub_weight_mom``, ``index``,
``num_local``, ``class_start``, ``num_sample``, ``stream``, file names), methods
(``forward_backward``, ``prepare``, ``sample``, ``update``, ``save_params``, ``save_FC``,
``update_FC``, ``update_from_tensor``) and optimizer contract.  What differs is underneath:

* the ``[Bt, Cs]`` logits are never materialised -- the margin and the softmax statistics live in the
  epilogue of a tcgen05 GEMM, and the backward recomputes them (``libfedfr_b200.so``);
* the three row-statistic all-reduces of partial_fc.py:142,147,161 are one all-gather of
  ``(max, sum-exp, target logit)`` triples followed by a local merge (mathematically identical);
* ``margin_softmax`` must be a ``CosFace(s, m)`` or ``ArcFace(s, m)`` object (this package's or the reference's); it
  is read as a descriptor, never called.

There is no PyTorch/CPU fallback: construction fails without an sm_100 device.
"""
import logging
import os

import torch
import torch.distributed as dist
from torch.nn import Module
from torch.nn.parameter import Parameter

from .losses import margin_params
from ._nvtx import nvtx_range


class PartialFC(Module):
    @torch.no_grad()
    def __init__(self, rank, local_rank, world_size, batch_size, resume, margin_softmax, num_classes, sample_rate=1.0,
                 embedding_size=512, prefix="./", check_mode=False, _ops=None):
        super(PartialFC, self).__init__()
        self.num_classes: int = num_classes
        self.rank: int = rank
        self.local_rank: int = local_rank
        self.world_size: int = world_size
        self.batch_size: int = batch_size
        self.margin_softmax = margin_softmax
        self._s, self._m, self._margin_kind = margin_params(margin_softmax)
        self.sample_rate: float = sample_rate
        self.embedding_size: int = embedding_size
        self.prefix: str = prefix
        # shard geometry, partial_fc.py:34-36
        self.num_local: int = num_classes // world_size + int(rank < num_classes % world_size)
        self.class_start: int = num_classes // world_size * rank + min(rank, num_classes % world_size)
        self.num_sample: int = int(self.sample_rate * self.num_local)

        if _ops is None:
            from . import _native as N
            from .ops_cuda import CudaOps
            self.device = torch.device("cuda:{}".format(self.local_rank))
            self._ops = CudaOps(self.device, N.PATH_CHECK if check_mode else N.PATH_TENSOR)
        else:                       # host-logic tests inject a CPU provider (tests/ only)
            self._ops = _ops
            self.device = torch.device(_ops.device)

        # checkpoint names, partial_fc.py:38-39
        self.weight_name = os.path.join(self.prefix, "rank:{}_softmax_weight.pt".format(self.rank))
        self.weight_mom_name = os.path.join(self.prefix, "rank:{}_softmax_weight_mom.pt".format(self.rank))

        if resume:      # partial_fc.py:41-54
            try:
                self.weight: torch.Tensor = torch.load(self.weight_name).to(self.device)
                logging.info("softmax weight resume successfully!")
            except (FileNotFoundError, KeyError, IndexError):
                self.weight = torch.normal(0, 0.01, (self.num_local, self.embedding_size), device=self.device)
                logging.info("softmax weight resume fail!")
            try:
                self.weight_mom: torch.Tensor = torch.load(self.weight_mom_name).to(self.device)
                logging.info("softmax weight mom resume successfully!")
            except (FileNotFoundError, KeyError, IndexError):
                self.weight_mom: torch.Tensor = torch.zeros_like(self.weight)
                logging.info("softmax weight mom resume fail!")
        else:           # partial_fc.py:55-60
            self.weight = torch.normal(0, 0.01, (self.num_local, self.embedding_size), device=self.device)
            self.weight_mom: torch.Tensor = torch.zeros_like(self.weight)
            logger = logging.getLogger('FL_face.partial')
            logger.info("softmax weight init successfully!")
            logger.info("softmax weight mom init successfully!")
        self.stream = torch.cuda.Stream(local_rank) if self.device.type == "cuda" else None

        self.index = None
        if int(self.sample_rate) == 1:      # partial_fc.py:64-67: the shard itself is the parameter
            self.update = lambda: 0
            self.sub_weight = Parameter(self.weight)
            self.sub_weight_mom = self.weight_mom
        else:
            self.sub_weight = Parameter(torch.empty((0, 0), device=self.device))
        self._norm = None       # (w_hat, inv_norm) of the current step
        self._range_checks = 0  # forward_backward calls seen by the range guard of the stored-probability path
        self._range_events = None
        self._range_slot = None
        self._prenorm = None    # (w_hat, inv_norm, weight ptr, weight version) left behind by step(prenormalize=True)
        self._label_buf = None
        self._bufs = {}
        self._flat_collectives = self._flat_rs = self._coalesce = None      # None = untried, True / False = backend support

    # ------------------------------------------------------------------ shard I/O (partial_fc.py:71-87)
    def save_params(self):
        torch.save(self.weight.data, self.weight_name)
        torch.save(self.weight_mom, self.weight_mom_name)

    def save_FC(self):
        torch.save(self.weight.data, os.path.join(self.prefix, 'FC_rank_%d.pth' % (self.local_rank)))

    def update_FC(self):
        model_path = os.path.join(self.prefix, 'FC_rank_%d.pth' % (self.local_rank))
        self.weight.data = torch.load(model_path).to(self.device)
        self.sub_weight = Parameter(self.weight)
        self._norm = self._prenorm = None       # normalised copies of the previous shard are stale
        print('Load weight from %s' % (model_path))

    def update_from_tensor(self, tensor):
        self.weight.data = tensor.to(self.device)
        self.sub_weight = Parameter(self.weight)
        self._norm = self._prenorm = None

    # ------------------------------------------------------------------ sampling (partial_fc.py:89-106)
    @torch.no_grad()
    def sample(self, total_label):
        """In place on ``total_label`` (global ids -> shard-local / sampled column ids, -1 elsewhere)."""
        ops = self._ops
        if hasattr(ops, "remap_labels_"):
            local = ops.remap_labels_(total_label, self.class_start, self.num_local)
        else:
            local = ops.remap_labels(total_label, self.class_start, self.num_local)
        if int(self.sample_rate) != 1:
            perm = torch.rand(size=[self.num_local], device=self.device)       # same draw as partial_fc.py:95
            index = ops.sample(local, perm, self.num_sample)
            self.index = index
            sub_w, sub_m = ops.gather_rows2(self.weight, self.weight_mom, index)
            self.sub_weight = Parameter(sub_w)
            self.sub_weight_mom = sub_m
        if local is not total_label:
            total_label.copy_(local)

    def forward(self, total_features, norm_weight):
        """The reference materialises ``logits = linear(total_features, norm_weight)`` here (partial_fc.py:108-111).
        The fused path never does; this method exists for API parity and returns them in fp32 for inspection."""
        return torch.nn.functional.linear(total_features, norm_weight.to(total_features.dtype))

    @torch.no_grad()
    def update(self):
        """Write the sampled rows back (partial_fc.py:113-116); replaced by a no-op when sample_rate == 1."""
        self._ops.scatter_rows2(self.weight, self.weight_mom, self.index, self.sub_weight.data, self.sub_weight_mom)

    @torch.no_grad()
    def step(self, optimizer, prenormalize=True):
        """``optimizer.step()`` + ``self.update()`` for the head in ONE pass over the shard (the step either
        side of the path).  Reads lr / momentum / dampening / weight_decay / nesterov from the optimizer's last param
        group -- the one ``prepare`` wired ``sub_weight`` and its momentum buffer into (partial_fc.py:124-126) -- and
        applies torch.optim.SGD's arithmetic in place to ``weight[index]`` / ``weight_mom[index]`` (so the scatter of
        partial_fc.py:113-116 never runs).  Use it INSTEAD of ``optimizer.step(); module.update()`` for an optimizer
        that holds only the head.  With ``sample_rate == 1`` and ``prenormalize`` the same pass also emits
        normalize(weight) of the updated shard, and the next ``forward_backward`` skips its normalisation pass."""
        group = optimizer.param_groups[-1]
        if len(optimizer.param_groups) != 1 or len(group["params"]) != 1 or group["params"][0] is not self.sub_weight:
            raise ValueError("PartialFC.step needs an optimizer whose only parameter is this module's sub_weight "
                             "(call forward_backward first; use optimizer.step() + update() otherwise)")
        if group.get("maximize", False) or float(group.get("momentum", 0.0)) == 0.0:
            raise NotImplementedError("PartialFC.step implements SGD with momentum (config.py:8), minimising")
        grad = self.sub_weight.grad
        if grad is None:
            return
        sampled = int(self.sample_rate) != 1
        pre = self._ops.sgd_step(self.weight, self.weight_mom, grad, self.index if sampled else None, group["lr"], group["momentum"],
                                 group.get("dampening", 0.0), group.get("weight_decay", 0.0), group.get("nesterov", False),
                                 prenormalize and not sampled)
        self._prenorm = None
        if pre is not None:
            self._prenorm = (pre[0], pre[1], self.weight.data_ptr(), self.weight._version)

    # ------------------------------------------------------------------ range guard of the stored-probability path
    _RANGE_LIMIT = 80.0         # nats: s * max|x_i| beyond this leaves the exponent window of include/fedfr_b200.h

    def _to_recompute(self, why):
        logging.getLogger('FL_face.partial').warning(
            "PartialFC: %s is outside the stored-probability range; using the recomputing backward from now on "
            "(normalise the embeddings, as partial_fc.py's callers do, to get the faster path)", why)
        self._ops.bwd_mode = "recompute"

    def _check_logit_range(self, features):
        """The stored-probability forward references every row to the bound s |x_i| (no row maximum is known before the
        GEMM); with un-normalised embeddings of large norm that bound is hundreds of nats above the real logits and they
        would underflow.  Two guards:
        * the FIRST call reads max |x_i| back (one small reduction, one host read) and decides before anything runs;
        * EVERY later step is checked on the device at no cost: its kernels raise a sticky flag pair in pinned host memory
          when s |x_i| leaves the window or a row sum comes out 0 (``pfc_set_range_flag``).  Step n reports into pair
          n mod 4 and the pair is read at the start of step n + 2, after an event recorded behind step n (long complete
          by then: no stall) -- every rank sees the same gathered batch, hence takes the same decision at the same step.
          A violation therefore switches the head to the recomputing backward two steps after it happened, with a log
          line (the offending rows of those steps had zero gradients), instead of going unnoticed for up to 255 steps."""
        ops = self._ops
        if getattr(ops, "bwd_mode", None) != "prob":
            return
        n = self._range_checks
        self._range_checks += 1
        if n == 0:
            worst = features.detach().to(torch.float32).norm(dim=1).max() * self._s
            if self.world_size > 1:
                dist.all_reduce(worst, op=dist.ReduceOp.MAX)
            if not float(worst) <= self._RANGE_LIMIT:          # also catches NaN
                self._to_recompute("s * |x| = %.1f" % float(worst))
                return
        if not hasattr(ops, "read_range_slot"):
            return
        if self._range_events is None:
            self._range_events = [torch.cuda.Event() for _ in range(4)]
        if n >= 2:
            self._range_events[(n - 2) % 4].synchronize()
            if ops.read_range_slot((n - 2) % 4):
                self._to_recompute("a batch two steps ago (s * |x| beyond %.0f nats, or an underflowed row)" % self._RANGE_LIMIT)
                return
        ops.set_range_slot(n % 4)
        self._range_slot = n % 4

    # ------------------------------------------------------------------ collectives
    def _buffer(self, name, shape, dtype):
        """Step scratch with a stable address (gather targets are arguments of the replayed forward / backward graphs)."""
        key = (name, tuple(shape), dtype)
        t = self._bufs.get(key)
        if t is None:
            t = torch.empty(tuple(shape), dtype=dtype, device=self.device)
            self._bufs[key] = t
        return t

    def _all_gather(self, out, inp):
        """out [W * n, ...] <- the W ranks' inp [n, ...] in rank order: ONE collective on the flat tensor (no per-rank chunk
        lists); backends without all_gather_into_tensor get the list form."""
        if self.world_size == 1:
            out.copy_(inp.reshape(out.shape))
            return
        if self._flat_collectives is not False:
            try:
                dist.all_gather_into_tensor(out, inp.contiguous())
                self._flat_collectives = True
                return
            except (RuntimeError, NotImplementedError, AttributeError):
                if self._flat_collectives:
                    raise
                self._flat_collectives = False
        dist.all_gather(list(out.chunk(self.world_size, dim=0)), inp)

    def _all_gather_pair(self, out_a, in_a, out_b, in_b):
        """The label gather (partial_fc.py:122) and the feature gather (:134) as ONE NCCL group launch: they have no
        dependence on each other and are latency bound at these sizes."""
        if self.world_size > 1 and self.device.type == "cuda" and self._coalesce is not False:
            try:
                with dist._coalescing_manager(device=self.device):
                    dist.all_gather_into_tensor(out_a, in_a)
                    dist.all_gather_into_tensor(out_b, in_b)
                self._coalesce = True
                return
            except (RuntimeError, NotImplementedError, AttributeError, ValueError):
                if self._coalesce:
                    raise
                self._coalesce = False
        self._all_gather(out_a, in_a)
        self._all_gather(out_b, in_b)

    def _reduce_scatter(self, out, inp):
        if self._flat_rs is not False:
            try:
                dist.reduce_scatter_tensor(out, inp)
                self._flat_rs = True
                return
            except (RuntimeError, NotImplementedError, AttributeError):
                if self._flat_rs:
                    raise
                self._flat_rs = False
        dist.reduce_scatter(out, list(inp.chunk(self.world_size, dim=0)))

    def _rewire(self, optimizer):
        """partial_fc.py:124-126: the optimizer's last param group follows ``sub_weight`` and its momentum buffer."""
        optimizer.state.pop(optimizer.param_groups[-1]['params'][0], None)
        optimizer.param_groups[-1]['params'][0] = self.sub_weight
        optimizer.state[self.sub_weight]['momentum_buffer'] = self.sub_weight_mom

    @torch.no_grad()
    def prepare(self, label, optimizer, _defer_normalize=False):
        """partial_fc.py:118-128: gather labels, sample, rewire the optimizer onto ``sub_weight`` and its
        momentum buffer, normalise the sub-shard.  Returns ``(total_label, norm_weight)``.
        (``forward_backward`` gathers labels and features together and defers the normalisation: it is fused with the
        logits kernel.)"""
        if self.world_size == 1:        # nothing to gather: the copy the reference's sample() mutates in place
            total_label = label.to(self.device, dtype=torch.long, copy=True)
        else:
            total_label = torch.zeros(size=[self.batch_size * self.world_size], device=self.device, dtype=torch.long)
            self._all_gather(total_label, label.to(self.device, dtype=torch.long))
        self.sample(total_label)
        self._rewire(optimizer)
        if _defer_normalize:
            return total_label, None
        self._norm = self._ops.normalize(self.sub_weight.data)
        return total_label, self._norm[0]

    @torch.no_grad()
    def forward_backward(self, label, features, optimizer):
        """partial_fc.py:130-176.  ``features`` fp32 [batch_size, E] (caller-normalised), ``label`` int64
        [batch_size].  Returns ``(x_grad [batch_size, E], loss)``; ``sub_weight.grad`` receives the shard gradient."""
        W, B, E = self.world_size, self.batch_size, self.embedding_size
        if features.shape[0] != B or label.shape[0] != B:
            raise ValueError("features/label batch must equal the constructor's batch_size (gather buffers are pre-sized, "
                             "partial_fc.py:120-122,132-134)")
        ops = self._ops
        fused = hasattr(ops, "normalize_fwd_stats")
        self._check_logit_range(features)
        w_hat = None
        nvtx_prep = nvtx_range("pfc.prepare(gather+sample)")
        nvtx_prep.__enter__()
        if W == 1:
            total_label, w_hat = self.prepare(label, optimizer, _defer_normalize=fused)
            if self._label_buf is None or self._label_buf.shape != total_label.shape:
                self._label_buf = torch.empty_like(total_label)
            self._label_buf.copy_(total_label)          # stable address for the replayed backward graph
            total_label = self._label_buf
            # nothing to gather (partial_fc.py:132-134 with one rank is a copy): read the features in place
            x_hat = ops.cast_features(features.data.to(device=self.device, dtype=torch.float32).contiguous())
        else:
            # partial_fc.py:120-122 and :132-134 in one group launch.  The features travel as the operand the kernels
            # consume (bf16 on the tensor path: cast(gather(x)) == gather(cast(x)), half the bytes), into step scratch
            x_loc = ops.cast_features(features.data.to(device=self.device, dtype=torch.float32).contiguous())
            total_label = self._buffer("total_label", (B * W,), torch.long)
            x_hat = self._buffer("x_hat_total", (B * W, E), x_loc.dtype)
            self._all_gather_pair(total_label, label.to(self.device, dtype=torch.long).contiguous(), x_hat, x_loc)
            self.sample(total_label)
            self._rewire(optimizer)
            if not fused:
                self._norm = ops.normalize(self.sub_weight.data)
                w_hat = self._norm[0]

        nvtx_prep.__exit__()
        # forward: per-shard (max, sum-exp, target logit), then one exchange instead of three all-reduces
        nvtx_fwd = nvtx_range("pfc.forward(normalize+logits+stats)")
        nvtx_fwd.__enter__()
        pre = self._prenorm
        self._prenorm = None
        if pre is not None and int(self.sample_rate) == 1 and pre[2] == self.weight.data_ptr() and pre[3] == self.weight._version:
            w_hat, inv_norm = pre[0], pre[1]        # step() already normalised the updated shard
            self._norm = (w_hat, inv_norm)
            stats = ops.fwd_stats(x_hat, w_hat, total_label, self._s, self._m, self._margin_kind)
        elif fused:     # normalize(sub_weight) chunk k+1 runs underneath the logits kernel of chunk k
            w_hat, inv_norm, stats = ops.normalize_fwd_stats(self.sub_weight.data, x_hat, total_label, self._s, self._m, self._margin_kind)
            self._norm = (w_hat, inv_norm)
        else:
            inv_norm = self._norm[1]
            stats = ops.fwd_stats(x_hat, w_hat, total_label, self._s, self._m, self._margin_kind)
        if W == 1:
            gathered = stats.unsqueeze(0)
        else:
            gathered = self._buffer("stats_all", (W,) + tuple(stats.shape), stats.dtype)
            self._all_gather(gathered.view(W * stats.shape[0], stats.shape[1]), stats)
        row_max, row_sum, loss_v = ops.finalize(gathered)
        nvtx_fwd.__exit__()

        # backward: dx partial of this shard + sub_weight.grad
        nvtx_bwd = nvtx_range("pfc.backward(dx+dw+reduce_scatter)")
        nvtx_bwd.__enter__()
        accumulate = self.sub_weight.grad is not None
        if not accumulate:
            self.sub_weight.grad = torch.empty_like(self.sub_weight.data)
        dx_total = ops.bwd(x_hat, w_hat, inv_norm, total_label, row_max, row_sum, self._s, self._m, 1.0 / (B * W),
                           self.sub_weight.grad, accumulate, self._margin_kind)

        if self._range_slot is not None:            # the range guard reads this step's flag pair two steps from now
            self._range_events[self._range_slot].record()
            self._range_slot = None
        if W == 1:
            x_grad = dx_total.clone()               # dx_total is step scratch with a stable address
        else:
            x_grad = torch.empty((B, E), dtype=torch.float32, device=self.device)
            self._reduce_scatter(x_grad, dx_total)  # partial_fc.py:171-173
            x_grad.mul_(W)                          # partial_fc.py:174
        nvtx_bwd.__exit__()
        return x_grad, loss_v
