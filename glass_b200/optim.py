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
        self._table_host = torch.zeros((len(self.params), 5), dtype=torch.int64).pin_memory()
        self._table_dev = torch.zeros((len(self.params), 5), dtype=torch.int64, device=dev)
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
        gradient every step, which is the case for the GLASS model).  The pointer table is rebuilt on
        every eager call; inside a CUDA-graph capture the (static) addresses are baked into the graph through a
        pinned staging copy."""
        rows = []
        for p, m, v in zip(self.params, self.exp_avg, self.exp_avg_sq):
            g = p.grad
            if g is None:
                rows.append((p.data_ptr(), 0, m.data_ptr(), v.data_ptr(), 0))
                continue
            if not g.is_contiguous() or g.dtype != torch.float32:
                raise RuntimeError("FusedAdam needs contiguous fp32 gradients")
            rows.append((p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel()))
        self._table_host.copy_(torch.tensor(rows, dtype=torch.int64))
        self._table_dev.copy_(self._table_host, non_blocking=True)
        lib = _lib.load()
        b1, b2 = self.betas
        check(lib.glass_adam_step(C.c_void_p(self._table_dev.data_ptr()), C.c_void_p(self.chunk_tensor.data_ptr()),
                                  C.c_void_p(self.chunk_begin.data_ptr()), self.chunk_tensor.numel(),
                                  C.c_void_p(self.lr.data_ptr()), C.c_void_p(self.state.data_ptr()), b1, b2, self.eps,
                                  self.weight_decay, C.c_void_p(torch.cuda.current_stream().cuda_stream)),
              "adam_step")
