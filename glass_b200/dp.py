"""Label-batch data parallelism over NVLink peer memory -- no library collective in the step.

`SymmetricGradExchange` places the trainable embedding table, its gradient, a block for all the small gradients and a
few flags in ONE symmetric (peer-mapped) allocation per rank (torch.distributed._symmetric_memory only provides the
allocation and the address exchange).  Per step, `step()` launches glass_dp_adam_step (csrc/dp.cu): reduce-scatter of
the table gradient, Adam on the owned rows, all-gather of the updated rows, and the average of the small gradients --
then the ordinary one-launch Adam updates the small parameters from the averaged block.

Numerics: the averaged gradient of every element is a sum in rank order divided by P, computed by exactly one rank
(table) or identically by every rank (small block), so all replicas hold bit-identical parameters after every step.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch

from . import _lib, ops
from ._lib import check
from .optim import FusedAdam

TABLE_MIN_ELEMS = 1 << 18        # parameters at least this large are exchanged as "the table" (row-partitioned)


def _align(v: int, a: int) -> int:
    return (v + a - 1) // a * a


def owned_range(n_elems: int, rank: int, world: int) -> tuple:
    """Element range of the (flattened) table that `rank` reduces and updates: contiguous, multiples of 4."""
    per = _align(-(-n_elems // world), 4)
    lo = min(rank * per, n_elems)
    return lo, min(lo + per, n_elems)


class SymmetricGradExchange:
    def __init__(self, params: List[torch.nn.Parameter], lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.params = [p for p in params if p.requires_grad]
        dev = self.params[0].device
        self.device = dev
        big = [p for p in self.params if p.numel() >= TABLE_MIN_ELEMS and p.numel() % 4 == 0]
        if len(big) > 1:
            raise NotImplementedError("one row-partitioned table per model (the GLASS models have one embedding table)")
        self.table: Optional[torch.nn.Parameter] = big[0] if big else None
        self.small = [p for p in self.params if p is not self.table]
        self.betas, self.eps, self.weight_decay = betas, eps, weight_decay
        lib = _lib.load()
        # ---- layout of the symmetric block (bytes)
        n_table = self.table.numel() if self.table is not None else 0
        n_small = _align(sum(p.numel() for p in self.small), 4)
        self.off_table = 0
        self.off_grad = _align(4 * n_table, 256)
        self.off_small = self.off_grad + _align(4 * n_table, 256)
        self.off_flags = self.off_small + _align(4 * n_small, 256)
        total = self.off_flags + _align(lib.glass_dp_flags_bytes(), 256)
        self.block = symm_mem.empty(total // 4, dtype=torch.float32, device=dev)
        self.block.zero_()
        self.handle = symm_mem.rendezvous(self.block, self.group)
        self.peer_base = torch.tensor([int(p) for p in self.handle.buffer_ptrs], dtype=torch.int64, device=dev)
        f32 = self.block
        self.n_table, self.n_small = n_table, n_small
        if self.table is not None:
            view = f32[self.off_table // 4:self.off_table // 4 + n_table].view_as(self.table)
            view.copy_(self.table.data)
            self.table.data = view                                   # the parameter now lives in peer-visible memory
            self.table_grad = f32[self.off_grad // 4:self.off_grad // 4 + n_table].view_as(self.table)
            ops.register_grad_buffer(self.table, self.table_grad)     # backward writes the table gradient here
            self.m = torch.zeros(n_table, dtype=torch.float32, device=dev)
            self.v = torch.zeros(n_table, dtype=torch.float32, device=dev)
            self.own = owned_range(n_table, self.rank, self.world)
        else:
            self.table_grad, self.m, self.v, self.own = None, None, None, (0, 0)
        # small gradients: staging views inside the symmetric block, averaged copy in local memory
        self.small_stage, self.small_avg_views, off = [], [], 0
        self.small_avg = torch.zeros(max(n_small, 4), dtype=torch.float32, device=dev)
        base = self.off_small // 4
        for p in self.small:
            self.small_stage.append(f32[base + off:base + off + p.numel()].view_as(p))
            self.small_avg_views.append(self.small_avg[off:off + p.numel()].view_as(p))
            off += p.numel()
        self.opt = FusedAdam(self.small, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay) if self.small else None
        self.lr = self.opt.lr if self.opt is not None else torch.tensor(float(lr), device=dev)
        self.state = self.opt.state if self.opt is not None else torch.zeros(2, dtype=torch.float32, device=dev)
        self.epoch = torch.zeros(1, dtype=torch.int64, device=dev)
        self.ticket = torch.zeros(1, dtype=torch.int32, device=dev)
        self.error = torch.zeros(1, dtype=torch.int32, device=dev)
        torch.cuda.synchronize(dev)
        dist.barrier(self.group)                                      # every rank's block is initialised

    def release(self):
        """Undo the constructor's side effects (table back in ordinary memory, gradient sink dropped)."""
        if self.table is not None:
            self.table.data = self.table.data.clone()
            ops.unregister_grad_buffer(self.table_grad)

    def zero_grad(self, set_to_none: bool = True):
        for p in self.params:
            p.grad = None

    def set_lr(self, lr: float):
        self.lr.fill_(float(lr))

    def reset_state(self):
        if self.opt is not None:
            self.opt.reset_state()
        else:
            self.state.zero_()
        if self.m is not None:
            self.m.zero_()
            self.v.zero_()

    @torch.no_grad()
    def step(self):
        """Exchange + optimizer for one training step (capturable in a CUDA graph)."""
        lib = _lib.load()
        if self.table is not None:
            g = self.table.grad
            if g is None:
                raise RuntimeError("the embedding table received no gradient this step")
            if g.data_ptr() != self.table_grad.data_ptr():            # not produced by the registered sink: copy in
                self.table_grad.copy_(g)
        pairs = [(s, p.grad) for s, p in zip(self.small_stage, self.small) if p.grad is not None]
        if pairs:
            torch._foreach_copy_([s for s, _ in pairs], [g for _, g in pairs])
        b1, b2 = self.betas
        vp = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        check(lib.glass_dp_adam_step(self.world, self.rank, vp(self.peer_base), self.off_table, self.off_grad,
                                     self.off_small, self.off_flags, self.n_table, self.own[0], self.own[1],
                                     self.n_small, vp(self.m), vp(self.v), vp(self.small_avg), vp(self.lr),
                                     vp(self.state), b1, b2, self.eps, self.weight_decay, vp(self.epoch),
                                     vp(self.ticket), vp(self.error),
                                     C.c_void_p(torch.cuda.current_stream().cuda_stream)), "dp_adam_step")
        ops._count(1)
        if self.opt is not None:
            for p, avg in zip(self.small, self.small_avg_views):       # the optimizer reads the averaged values
                if p.grad is not None:
                    p.grad = avg
            self.opt.step()
        elif self.table is not None:                                   # nobody else advances the step count
            self.state[0] += 1

    def check_error(self):
        """Raise if a peer failed to answer inside the kernel's time-out (one host sync; call between epochs)."""
        if int(self.error.item()) != 0:
            raise RuntimeError("glass_dp_adam_step: a peer rank did not reach the gradient exchange (time-out)")
