"""Whole-step CUDA graph: labels -> forward -> loss -> backward -> (all-reduce) -> Adam, one replay per step.

The reference's loop (impl/train.py:8-16) issues ~10^2 small launches and one host sync per step; on
every shipped config the GPU work per step is shorter than the Python time needed to launch it
(SURVEY.md section 7 "hard parts" 1).  All kernels of libglass_b200.so are capture-safe (no allocation,
no host sync), shapes are static per dataset (fixed batch size with drop_last, globally padded
subG_node), so the step is captured once and replayed with new (subG_node, y) copied into static
buffers.  Semantics are those of impl/train.py:10-16 with Adam (GLASSTest.py:213).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from . import utils


class GradAverager:
    """Averages gradients across ranks with as few NCCL launches as possible: every large tensor (the
    trainable N x H embedding table is ~99 % of the bytes) is reduced in place, all small ones travel
    through one persistent flat buffer (a coalesced group of 17 separate all-reduces costs ~100 us of
    per-operation latency at N = 2; one 14.8 MB all-reduce costs 55 us).  Capturable in a CUDA graph."""

    BIG = 1 << 18   # elements

    def __init__(self, params, group=None):
        self.group = group
        self.small = [p for p in params if p.numel() < self.BIG]
        self.big = [p for p in params if p.numel() >= self.BIG]
        total = sum(p.numel() for p in self.small)
        ref = params[0]
        self.flat = torch.zeros(max(total, 1), dtype=ref.dtype, device=ref.device)
        self.views, off = [], 0
        for p in self.small:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    def _average(self, t):
        import torch.distributed as dist
        if dist.get_backend(self.group) == "nccl":
            dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group)
        else:                       # gloo (CPU tests of the host logic) has no AVG
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            t.div_(dist.get_world_size(self.group))

    def __call__(self):
        for p in self.big:
            if p.grad is not None:
                self._average(p.grad)
        pairs = [(v, p) for v, p in zip(self.views, self.small) if p.grad is not None]
        if pairs:
            torch._foreach_copy_([v for v, _ in pairs], [p.grad for _, p in pairs])
            self._average(self.flat)
            for v, p in pairs:      # the optimizer reads the averaged values straight from the flat buffer
                p.grad = v


class GraphedTrainStep:
    def __init__(self, model, loss_fn: Callable, x, edge_index, edge_weight, pos_example: torch.Tensor,
                 y_example: torch.Tensor, lr: float, group=None, warmup: int = 3, betas=(0.9, 0.999), eps=1e-8,
                 weight_decay: float = 0.0, z_fn=utils.MaxZOZ):
        dev = x.device
        self.model, self.loss_fn, self.z_fn = model, loss_fn, z_fn
        self.x, self.ei, self.ew = x, edge_index, edge_weight
        self.pos = torch.empty_like(pos_example, device=dev)
        self.y = torch.empty_like(y_example, device=dev)
        self.pos.copy_(pos_example)
        self.y.copy_(y_example)
        import torch.distributed as dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        # autograd hands each parameter its freshly written gradient tensor (no zero-fill, no accumulate pass).
        # Several ranks: glass_b200.dp.SymmetricGradExchange (one fused reduce-scatter + Adam + all-gather launch over
        # NVLink peer memory); GLASS_B200_DP=nccl keeps the round-1 path (NCCL all-reduce, then Adam).
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.averager = None
        from .optim import FusedAdam
        import os
        self.dp_mode = "none"
        if self.world > 1 and os.environ.get("GLASS_B200_DP", "symm") == "symm":
            opt, why = None, ""
            try:
                from .dp import SymmetricGradExchange
                opt = SymmetricGradExchange(self.params, lr, betas, eps, weight_decay, group)
            except (NotImplementedError, ImportError, AttributeError, RuntimeError) as e:
                why = str(e)
            # the choice is collective: one rank without peer-mapped memory sends every rank to the NCCL path
            ok = torch.tensor([1 if opt is not None else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 1:
                self.opt, self.dp_mode = opt, "symm"
            else:
                import warnings
                warnings.warn(f"symmetric-memory gradient exchange unavailable ({why or 'on a peer rank'}); using NCCL all-reduce")
                if opt is not None:
                    opt.release()
        if self.dp_mode == "none":
            if self.world > 1:
                self.averager = GradAverager(self.params, group)
                self.dp_mode = "nccl"
            self.opt = FusedAdam(model.parameters(), lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self.lr = self.opt.lr
        self.loss = torch.zeros((), device=dev)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        model.train()
        # warm-up on a side stream (builds the CSR cache, initialises Adam state, primes the allocator)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._step_eager()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)

    # one iteration of impl/train.py:10-16
    def _step_eager(self):
        z = self.z_fn(self.x, self.pos)
        self.opt.zero_grad(set_to_none=True)
        loss = self.loss_fn(self.model(self.x, self.ei, self.ew, self.pos, z, id=0), self.y)
        loss.backward()
        if self.averager is not None:
            self.averager()
        self.opt.step()
        self.loss.copy_(loss.detach())

    def capture(self):
        """Capture one training step.  NOTE: the warm-up steps in __init__ already updated the parameters and
        the Adam state; callers that need the pre-warm-up model snapshot its state_dict before constructing
        this object and call reset_to(snapshot) after capture()."""
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._step_eager()
        return self

    def reset_to(self, state_dict):
        """Restore parameters and clear Adam moments / step counters (in place: graph addresses stay valid)."""
        with torch.no_grad():
            own = self.model.state_dict()
            for k, v in state_dict.items():
                own[k].copy_(v)
            self.opt.reset_state()

    def set_lr(self, lr: float):
        self.lr.fill_(float(lr))

    def __call__(self, pos: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        """pos/y may live on the host (pinned) or the device; returns the (device) loss of this step."""
        self.pos.copy_(pos, non_blocking=True)
        self.y.copy_(y, non_blocking=True)
        if self.graph is None:
            self._step_eager()
        else:
            self.graph.replay()
        return self.loss


class GraphedForward:
    """Inference pass of impl/train.py:28-29 (labels + model.eval() forward) as one CUDA graph."""

    def __init__(self, model, x, edge_index, edge_weight, pos_example: torch.Tensor, z_fn=utils.MaxZOZ, warmup: int = 2):
        dev = x.device
        self.model, self.x, self.ei, self.ew, self.z_fn = model, x, edge_index, edge_weight, z_fn
        self.pos = torch.empty_like(pos_example, device=dev)
        self.pos.copy_(pos_example)
        model.eval()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):
                self.out = self.model(self.x, self.ei, self.ew, self.pos, self.z_fn(self.x, self.pos))
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.out = self.model(self.x, self.ei, self.ew, self.pos, self.z_fn(self.x, self.pos))

    def __call__(self, pos: torch.Tensor) -> torch.Tensor:
        self.pos.copy_(pos, non_blocking=True)
        self.graph.replay()
        return self.out


class GraphedSharedBaseForward:
    """Evaluation over MANY label batches (impl/train.py:20-34) with everything that does not depend on the label
    batch computed once per epoch: normalised input embedding, first-layer transform of
    both branches, and the dense adj @ U.  Per batch one CUDA-graph replay runs the sparse label correction
    (ops.spmm_delta), the combine GEMM, deeper layers if any, pooling and the head.  Same logits as GraphedForward
    up to fp32 re-association (adj @ U + correction instead of one sum).

    refresh() must be called after the weights change (once per evaluation epoch); it rewrites the shared buffers
    in place, so the captured per-batch graph stays valid."""

    def __init__(self, model, x, edge_index, edge_weight, pos_example: torch.Tensor, z_fn=utils.MaxZOZ, warmup: int = 2):
        dev = x.device
        self.model, self.x, self.ei, self.ew, self.z_fn = model, x, edge_index, edge_weight, z_fn
        self.pos = torch.empty_like(pos_example, device=dev)
        self.pos.copy_(pos_example)
        model.eval()
        self.base = None
        self.base_graph = None
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            self.base = tuple(t.clone() for t in model.shared_base(x, edge_index, edge_weight))   # static buffers
            for _ in range(warmup):
                self._refresh_eager()
                self.out = self._batch()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.base_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.base_graph), torch.no_grad():
            self._refresh_eager()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.out = self._batch()

    def _refresh_eager(self):
        fresh = self.model.shared_base(self.x, self.ei, self.ew)
        for dst, src in zip(self.base, fresh):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src)

    def _batch(self):
        return self.model.forward_from_base(self.base, self.ei, self.ew, self.pos, self.z_fn(self.x, self.pos))

    def refresh(self):
        """Recompute the label-independent base for the model's CURRENT weights (one graph replay)."""
        self.model.eval()
        self.base_graph.replay()

    def __call__(self, pos: torch.Tensor) -> torch.Tensor:
        self.pos.copy_(pos, non_blocking=True)
        self.graph.replay()
        return self.out


def train_epoch(step: GraphedTrainStep, batches, sync_each_step: bool = True) -> float:
    """impl/train.py:4-17 over an iterable of (subG_node, y) batches using the captured step."""
    losses = []
    for pos, y in batches:
        loss = step(pos, y)
        losses.append(loss.item() if sync_each_step else loss.clone())
    if sync_each_step:
        return sum(losses) / len(losses)
    return float(torch.stack(losses).double().sum().item()) / len(losses)


@torch.no_grad()
def test_epoch(fwd: GraphedForward, loader, metrics, loss_fn):
    """impl/train.py:20-34 with the captured forward.  A ragged last batch (drop_last=False loaders) is
    padded with empty subgraphs (rows of -1): they add no labels, pool to zero, and their logits are dropped,
    so the result equals train.test on the same loader."""
    from .SubGDataset import epoch_batches
    cap = fwd.pos.shape[0]
    if hasattr(fwd, "refresh"):      # shared-base evaluator: the weights may have changed since the last epoch
        fwd.refresh()
    preds, ys = [], []
    for pos, y in epoch_batches(loader):
        n = pos.shape[0]
        if n > cap:
            raise RuntimeError(f"batch of {n} subgraphs exceeds the captured batch size {cap}")
        if n < cap:
            pos = torch.cat((pos, pos.new_full((cap - n, pos.shape[1]), -1)), dim=0)
        preds.append(fwd(pos)[:n].clone())
        ys.append(y)
    pred, y = torch.cat(preds, dim=0), torch.cat(ys, dim=0)
    return metrics(pred.cpu().numpy(), y.cpu().numpy()), loss_fn(pred, y)
