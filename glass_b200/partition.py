"""Row-partitioned SpMM for graphs that should not be replicated (stress config:
2M nodes / 100M undirected edges).  New capability -- the reference keeps the whole graph on one device.

Rank r owns a contiguous block of rows of A (and of A^T), chosen so that every rank holds about the same
number of stored entries (power-law graphs: balance nnz, not rows), plus the matching rows of every dense
per-node matrix.  One exchange step per SpMM:

    forward : all-gather the owners' feature rows -> local block SpMM        y_r  = A[rows_r, :]   @ X
    backward: all-gather the owners' dY rows      -> local block SpMM        dX_r = A^T[rows_r, :] @ dY

Feature shards are stored padded to the largest shard so that one `all_gather_into_tensor` (NCCL over
NVLink) moves everything; the column indices of the local CSR blocks are remapped once to that padded
layout.  GEMMs, label mix, pooling gathers are row-local and need no communication; GraphNorm needs one
2C-value all-reduce of the column sums (not wired into the modules yet).
"""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist

from . import ops


def balanced_row_splits(rowptr: torch.Tensor, parts: int) -> List[int]:
    """Row boundaries [0 = b_0 <= b_1 <= ... <= b_P = N] with about nnz/P stored entries per block."""
    n = rowptr.numel() - 1
    nnz = int(rowptr[-1])
    targets = torch.arange(1, parts, dtype=torch.float64) * (nnz / parts)
    cuts = torch.searchsorted(rowptr.to(torch.float64).cpu(), targets).clamp(0, n).tolist()
    bounds = [0] + [int(c) for c in cuts] + [n]
    for i in range(1, len(bounds)):
        bounds[i] = max(bounds[i], bounds[i - 1])
    return bounds


def _block(rowptr, col, val, lo, hi, remap):
    s, e = int(rowptr[lo]), int(rowptr[hi])
    rp = (rowptr[lo:hi + 1] - rowptr[lo]).to(torch.int32).contiguous()
    c = remap[col[s:e].long()].to(torch.int32).contiguous()
    return rp, c, val[s:e].contiguous()


def _split_columns(rp, c, v, lo_col: int, hi_col: int):
    """Split a CSR block into the entries whose column lies in [lo_col, hi_col) -- re-based to 0, they index
    this rank's own feature shard -- and the rest (columns unchanged).  Entry order inside a row is kept."""
    n_rows = rp.numel() - 1
    counts = (rp[1:] - rp[:-1]).long()
    rows = torch.repeat_interleave(torch.arange(n_rows, device=c.device), counts)
    own = (c >= lo_col) & (c < hi_col)

    def pick(sel, shift):
        cnt = torch.bincount(rows[sel], minlength=n_rows)
        rp_new = torch.zeros(n_rows + 1, dtype=torch.int64, device=c.device)
        torch.cumsum(cnt, 0, out=rp_new[1:])
        return rp_new.to(torch.int32), (c[sel] - shift).contiguous(), v[sel].contiguous()

    return pick(own, lo_col), pick(~own, 0)


class RowPartitionedAdj:
    """This rank's row blocks of A and A^T with columns remapped to the padded all-gather layout.

    overlap=True additionally splits each block by column ownership: the entries that touch this rank's own
    feature rows are multiplied while the all-gather is in flight, the remote ones afterwards (SURVEY.md
    section 8e "start SpMM on locally-owned columns first").  The two partial products are added, so the
    summation order differs from the replicated SpMM (last-bit differences); overlap=False is bit-identical.
    A kernel that gathers straight from peer memory instead is the wrong trade here: peer loads bypass the
    local L2, so every remote entry would pull its 256-byte row over NVLink (12.8 GB per SpMM at N = 2 on the
    stress graph) against 256 MB for the all-gather."""

    def __init__(self, adj: ops.CSRAdj, rank: int, world: int, group=None, overlap: bool = False):
        self.rank, self.world, self.group, self.n = rank, world, group, adj.n
        dev = adj.col.device
        self.bounds = balanced_row_splits(adj.rowptr, world)
        sizes = [self.bounds[i + 1] - self.bounds[i] for i in range(world)]
        self.rows = sizes[rank]
        self.pad = max(max(sizes), 1)                       # rows per shard in the gathered buffer
        self.lo, self.hi = self.bounds[rank], self.bounds[rank + 1]
        # global row id -> row in the [world * pad, H] gathered matrix
        owner = torch.bucketize(torch.arange(adj.n, device=dev), torch.tensor(self.bounds[1:-1], device=dev), right=True)
        starts = torch.tensor(self.bounds[:-1], device=dev)
        remap = owner * self.pad + (torch.arange(adj.n, device=dev) - starts[owner])
        blk = _block(adj.rowptr, adj.col, adj.val, self.lo, self.hi, remap)
        blk_t = _block(adj.rowptr_t, adj.col_t, adj.val_t, self.lo, self.hi, remap)
        none = None
        self.local = ops.CSRAdj(self.rows, *blk, *blk_t, none, adj.aggr)
        self.local.make_plans()
        self.nnz_local = int(blk[1].numel())
        self.gather_override = None     # tests: callable(padded_local) -> [world * pad, H] without a process group
        self.own = self.rem = None
        if overlap:
            lo_col = rank * self.pad
            (o, r), (ot, rt) = _split_columns(*blk, lo_col, lo_col + self.rows), _split_columns(*blk_t, lo_col, lo_col + self.rows)
            self.own = ops.CSRAdj(self.rows, *o, *ot, none, adj.aggr).make_plans()
            self.rem = ops.CSRAdj(self.rows, *r, *rt, none, adj.aggr).make_plans()

    def shard(self, full: torch.Tensor) -> torch.Tensor:
        """Rows of a replicated [N, H] matrix owned by this rank."""
        return full[self.lo:self.hi].contiguous()

    def _gather_async(self, local: torch.Tensor):
        """Start the all-gather; returns (gathered buffer, wait()) -- wait() orders the current stream after it."""
        h = local.shape[1]
        send = local
        if self.rows != self.pad:
            send = torch.zeros((self.pad, h), dtype=local.dtype, device=local.device)
            send[:self.rows] = local
        send = send.contiguous()
        if self.gather_override is not None:
            out = self.gather_override(send)
            return out, (lambda: None)
        out = torch.empty((self.world * self.pad, h), dtype=local.dtype, device=local.device)
        if self.world > 1:
            work = dist.all_gather_into_tensor(out, send, group=self.group, async_op=True)
            return out, work.wait
        out.copy_(send)
        return out, (lambda: None)

    def _spmm_overlapped(self, v_local: torch.Tensor, transposed: bool) -> torch.Tensor:
        full, wait = self._gather_async(v_local)
        own, rem = (self.own.t(), self.rem.t()) if transposed else (self.own, self.rem)
        y = torch.empty((self.rows, v_local.shape[1]), dtype=torch.float32, device=v_local.device)
        ops._run_spmm(own.rowptr, own.col, own.val, own.plan, v_local.contiguous(), y)      # overlaps the gather
        wait()
        y2 = torch.empty_like(y)
        ops._run_spmm(rem.rowptr, rem.col, rem.val, rem.plan, full, y2)
        return y.add_(y2)

    def _gather(self, local: torch.Tensor) -> torch.Tensor:
        out, wait = self._gather_async(local)
        wait()
        return out

    def spmm(self, x_local: torch.Tensor) -> torch.Tensor:
        """y_local = A[rows_r, :] @ X with X assembled from every rank's shard (differentiable)."""
        return _PartSpMM.apply(x_local, self)


class _PartSpMM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_local, part: RowPartitionedAdj):
        ctx.part = part
        if part.own is not None:
            return part._spmm_overlapped(x_local, False)
        x_full = part._gather(x_local)
        y = torch.empty((part.rows, x_local.shape[1]), dtype=torch.float32, device=x_local.device)
        a = part.local
        ops._run_spmm(a.rowptr, a.col, a.val, a.plan, x_full, y)
        return y

    @staticmethod
    def backward(ctx, gy):
        part = ctx.part
        if part.own is not None:
            return part._spmm_overlapped(gy.contiguous(), True), None
        gy_full = part._gather(gy.contiguous())
        gx = torch.empty((part.rows, gy.shape[1]), dtype=torch.float32, device=gy.device)
        a = part.local
        ops._run_spmm(a.rowptr_t, a.col_t, a.val_t, a.plan_t, gy_full, gx)
        return gx, None
