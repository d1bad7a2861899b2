"""Adam as ONE kernel launch over all parameters (glass_adam_step), for the captured train step.

Same update rule and defaults as ``torch.optim.Adam`` (GLASSTest.py:213); the learning rate and the step count
live in device memory so that a CUDA graph containing the step can be replayed and the scheduler can change
the learning rate between replays (``set_lr``)."""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List

import torch

from . import _lib
from ._lib import check


class FusedAdam:
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FusedAdam: no trainable parameters")
        dev = self.params[0].device
        for p in self.params:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous() or p.device != dev:
                raise RuntimeError("FusedAdam needs contiguous fp32 CUDA parameters on one device")
        self.betas, self.eps, self.weight_decay = betas, eps, weight_decay
        self.lr = torch.tensor(float(lr), device=dev)
        self.state = torch.zeros(2, dtype=torch.float32, device=dev)          # {step count, ticket}
        self.exp_avg = [torch.zeros_like(p) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p) for p in self.params]
        chunk = _lib.load().glass_adam_chunk()
        ct, cb = [], []
        for i, p in enumerate(self.params):
            for b in range(0, p.numel(), chunk):
                ct.append(i)
                cb.append(b)
        self.chunk_tensor = torch.tensor(ct, dtype=torch.int32, device=dev)
        self.chunk_begin = torch.tensor(cb, dtype=torch.int64, device=dev)
        # Pointer table {p, g, m, v, n} per tensor.  Eager steps: a ring of pinned staging buffers, each guarded by
        # an event (the CPU may run ahead of the GPU; a queued H2D copy must never see a rewritten buffer) and an
        # upload only when an address changed.  Captured steps: a device table and a pinned buffer of their OWN
        # (allocated here, never rewritten after the capture), so later eager steps cannot corrupt a graph.
        shape = (len(self.params), 5)
        self._stage = [(torch.zeros(shape, dtype=torch.int64).pin_memory(), torch.cuda.Event()) for _ in range(4)]
        self._stage_i = 0
        self._last_rows = None
        self._table_dev = torch.zeros(shape, dtype=torch.int64, device=dev)
        self._capture_pool = [(torch.zeros(shape, dtype=torch.int64).pin_memory(),
                               torch.zeros(shape, dtype=torch.int64, device=dev)) for _ in range(4)]
        self._captured = []
        self.device = dev

    def zero_grad(self, set_to_none: bool = True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def set_lr(self, lr: float):
        self.lr.fill_(float(lr))

    def reset_state(self):
        self.state.zero_()
        for t in self.exp_avg + self.exp_avg_sq:
            t.zero_()

    @torch.no_grad()
    def step(self):
        """Parameters without a gradient this step are skipped (like torch; the step count used for the bias
        correction is global here, per parameter in torch -- identical as long as every parameter receives a
        gradient every step, which is the case for the GLASS model)."""
        rows = []
        for p, m, v in zip(self.params, self.exp_avg, self.exp_avg_sq):
            g = p.grad
            if g is None:
                rows.append((p.data_ptr(), 0, m.data_ptr(), v.data_ptr(), 0))
                continue
            if not g.is_contiguous() or g.dtype != torch.float32:
                raise RuntimeError("FusedAdam needs contiguous fp32 gradients")
            rows.append((p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel()))
        if torch.cuda.is_current_stream_capturing():
            if not self._capture_pool:
                raise RuntimeError("FusedAdam: more than 4 captures of one optimizer")
            host, table = self._capture_pool.pop()
            host.copy_(torch.tensor(rows, dtype=torch.int64))
            table.copy_(host, non_blocking=True)       # memcpy node: replays re-read `host`, which is never rewritten
            self._captured.append((host, table))
        else:
            table = self._table_dev
            if rows != self._last_rows:
                host, ev = self._stage[self._stage_i]
                ev.synchronize()                        # the copy that last used this buffer has completed
                host.copy_(torch.tensor(rows, dtype=torch.int64))
                table.copy_(host, non_blocking=True)
                ev.record()
                self._stage_i = (self._stage_i + 1) % len(self._stage)
                self._last_rows = rows
        lib = _lib.load()
        b1, b2 = self.betas
        check(lib.glass_adam_step(C.c_void_p(table.data_ptr()), C.c_void_p(self.chunk_tensor.data_ptr()),
                                  C.c_void_p(self.chunk_begin.data_ptr()), self.chunk_tensor.numel(),
                                  C.c_void_p(self.lr.data_ptr()), C.c_void_p(self.state.data_ptr()), b1, b2, self.eps,
                                  self.weight_decay, C.c_void_p(torch.cuda.current_stream().cuda_stream)),
              "adam_step")
