"""Fused native train / inference steps (the fast path behind the Lightning-shaped modules).

One step = native forward (activations kept) -> fused loss head (value + dOut) -> native backward
(flat gradient buffer) -> [data parallel: ONE all-reduce of the flat gradient buffer over NCCL]
-> fused Adam / SGD on the flat buffers.  Autograd is not involved; the parameters the user sees
(``module.parameters()``) are views of the flat buffer, so they are updated in place.

Data parallelism: graphs are independent, parameters are small and replicated; each
rank runs its shard of the graph batch, the loss gradient is pre-scaled by 1/world_size and the
flat fp32 gradient buffer (8.6 MB for the K4 model) is summed with a single all-reduce.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _native as N


def allreduce_flat_(grads: torch.Tensor, world: int, group=None) -> torch.Tensor:
    """The one collective of the data-parallel step: sum the flat gradient buffer over the ranks.

    Each rank's loss gradient is pre-scaled by 1/world (``loss_scale``), so the sum equals the gradient of the
    mean loss over the global batch when the shards have equal size; no post-divide is needed."""
    if world > 1:
        torch.distributed.all_reduce(grads, op=torch.distributed.ReduceOp.SUM, group=group)
    return grads


class FusedTrainer:
    def __init__(self, module, optimizer: Optional[str] = None, lr: Optional[float] = None, process_group=None,
                 betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        """``module``: a Lightning-shaped module (``.model``, ``.regression``, ``.optimizer``, ``.lr``) or a bare
        native model (then pass ``optimizer`` / ``lr``)."""
        self.module = module
        self.model = getattr(module, "model", module)
        self.optimizer = optimizer or getattr(module, "optimizer", "adam")
        self.lr = float(lr if lr is not None else getattr(module, "lr", 1e-3))
        if self.optimizer not in ("adam", "sgd"):
            raise ValueError("Invalid optimizer setting")
        self.betas, self.eps, self.weight_decay = betas, eps, weight_decay
        self.pg = process_group
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        self.step_count = 0
        self.grads = None
        self.exp_avg = None
        self.exp_avg_sq = None
        # data parallel: the layer-stack gradients (everything after the encoder block of the flat buffer) are all-reduced on
        # a side stream underneath the encoder weight gradient; `overlap_allreduce = False` keeps one all-reduce at the end
        self.overlap_allreduce = True
        self.skip_allreduce = False          # measurement hook of bench.py / the tests (isolates the collective's cost); never set by the library
        self._side = None
        self._layers_ready = None
        self._synced_params = False
        self._graphs = {}

    def _loss_kind(self) -> int:
        regression = getattr(self.module, "regression", getattr(self.model, "regression", True))
        return N.LOSS_MSE if regression else N.LOSS_CE2

    def _prepare(self, batch):
        x_dict = batch.x_dict
        model = self.model
        eng, B = model._engine_for(x_dict, batch.edge_index_dict)
        model._last_engine = eng
        dev = x_dict[model.node_types[0]].device
        model._ensure_flat(dev)
        return eng, B, [x_dict[t] for t in model.node_types], dev

    def _sync_parameters(self, flat: torch.Tensor) -> None:
        """Replicas must start from the same weights: rank 0's flat buffer is broadcast once, before the first step
        (DDP does the same at construction)."""
        if self.world > 1 and not self._synced_params:
            torch.distributed.broadcast(flat, src=torch.distributed.get_global_rank(self.pg, 0) if self.pg is not None else 0, group=self.pg)
        self._synced_params = True

    def _backward_and_reduce(self, eng, dout, flat, dev) -> None:
        """Native backward, then the gradient all-reduce: in two buckets when there is more than one rank - everything after
        the encoder block (7.5 of 8.6 MB for the K4 model) as soon as the layer-stack reduction has run, on a side stream
        underneath the encoder weight gradient; the encoder block at the end."""
        if self.world == 1 or self.skip_allreduce or not self.overlap_allreduce:
            eng.backward(dout, flat, grads=self.grads)
            if not self.skip_allreduce:
                allreduce_flat_(self.grads, self.world, self.pg)
            return
        if self._side is None:
            self._side = torch.cuda.Stream(dev)
            self._layers_ready = torch.cuda.Event()
        cur = torch.cuda.current_stream(dev)
        eng.backward(dout, flat, grads=self.grads, layers_ready=self._layers_ready)
        cut = eng.encoder_param_end()
        self._side.wait_event(self._layers_ready)
        with torch.cuda.stream(self._side):
            torch.distributed.all_reduce(self.grads[cut:], op=torch.distributed.ReduceOp.SUM, group=self.pg)
        torch.distributed.all_reduce(self.grads[:cut], op=torch.distributed.ReduceOp.SUM, group=self.pg)
        cur.wait_stream(self._side)

    def train_step(self, batch) -> torch.Tensor:
        """Runs one optimisation step on ``batch`` (device tensors); returns the loss as a [1] device tensor."""
        eng, B, xs, dev = self._prepare(batch)
        model = self.model
        flat = model._flat
        n = flat.numel()
        if self.grads is None or self.grads.device != dev or self.grads.numel() != n:
            self.grads = torch.empty(n, dtype=torch.float32, device=dev)
            self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
            self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self._sync_parameters(flat)
        out = eng.forward(xs, flat, train=True)
        model._fwd_token = object()
        loss, dout = eng.loss(out, batch.y, self._loss_kind(), want_grad=True, loss_scale=1.0 / self.world)
        self._backward_and_reduce(eng, dout, flat, dev)
        self.step_count += 1
        self._optimizer_step(flat, n, dev)
        return loss

    def _optimizer_step(self, flat, n, dev) -> None:
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            if self.optimizer == "adam":
                N.adam_step(flat.data_ptr(), self.grads.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), n,
                            self.step_count, self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, stream)
            else:
                N.sgd_step(flat.data_ptr(), self.grads.data_ptr(), n, self.lr, stream)

    # ---- CUDA graph of the forward + loss + backward of a FIXED batch object (small batches are launch bound) ----------
    def train_step_graphed(self, batch) -> torch.Tensor:
        """Same result as ``train_step`` for a batch whose tensors keep their addresses between calls (e.g. a staging
        batch the caller copies new windows into): forward, loss head and backward (the ~20 native launches) are replayed
        from one CUDA graph, the gradient all-reduce and the optimizer (whose bias correction changes per step) are
        launched normally.  The graph is captured at the first call for this batch object, after one eager step."""
        key = id(batch)
        ent = self._graphs.get(key)
        if ent is None:
            loss = self.train_step(batch)                 # warm-up: plan upload, workspace, kernel attributes, NCCL
            eng, B, xs, dev = self._prepare(batch)
            model = self.model
            flat = model._flat
            g = torch.cuda.CUDAGraph()
            static_loss = torch.empty(1, dtype=torch.float32, device=dev)
            torch.cuda.synchronize(dev)
            with torch.cuda.graph(g):
                out = eng.forward(xs, flat, train=True)
                l, dout = eng.loss(out, batch.y, self._loss_kind(), want_grad=True, loss_scale=1.0 / self.world)
                eng.backward(dout, flat, grads=self.grads)
                static_loss.copy_(l)
            self._graphs[key] = (g, static_loss, batch, flat.numel(), dev)     # holds the batch: its addresses stay valid
            return loss
        g, static_loss, _, n, dev = ent
        model = self.model
        flat = model._flat
        g.replay()
        model._fwd_token = object()
        if not self.skip_allreduce:
            allreduce_flat_(self.grads, self.world, self.pg)
        self.step_count += 1
        self._optimizer_step(flat, n, dev)
        return static_loss

    @torch.no_grad()
    def infer(self, batch) -> torch.Tensor:
        """No-grad forward through the native engine (ping-pong activation buffers, nothing saved)."""
        eng, B, xs, dev = self._prepare(batch)
        return self.model._finish(eng.forward(xs, self.model._flat, train=False), B)
