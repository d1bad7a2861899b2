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

import os
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


def split_columns_by_owner(rp, c, v, bounds, rebase: bool = True):
    """Split a CSR block by column range: part s holds the entries with bounds[s] <= column < bounds[s + 1] (columns
    re-based to bounds[s] when `rebase`: they then index owner s's own feature shard).  Entry order inside a row is
    kept.  Returns [(rowptr int32, col int32, val)] * (len(bounds) - 1)."""
    n_rows = rp.numel() - 1
    counts = (rp[1:] - rp[:-1]).long()
    rows = torch.repeat_interleave(torch.arange(n_rows, device=c.device), counts)
    edges = torch.tensor(list(bounds[1:-1]), dtype=c.dtype, device=c.device)
    owner = torch.bucketize(c, edges, right=True)
    out = []
    for s_ in range(len(bounds) - 1):
        sel = owner == s_
        cnt = torch.bincount(rows[sel], minlength=n_rows)
        rp_new = torch.zeros(n_rows + 1, dtype=torch.int64, device=c.device)
        torch.cumsum(cnt, 0, out=rp_new[1:])
        shift = int(bounds[s_]) if rebase else 0
        out.append((rp_new.to(torch.int32), (c[sel] - shift).to(torch.int32).contiguous(), v[sel].contiguous()))
    return out


MAX_PHASES = int(os.environ.get("GLASS_B200_PARTITION_PHASES", "4"))


def phase_groups(rank: int, world: int, max_phases: int = None):
    """Source ranks per phase of the pipelined product: [[rank], [rank+1, ...], ...] -- the own shard first, then the
    peers in arrival order cut into at most max_phases - 1 groups of consecutive sources (every phase walks the row
    structure and re-reads y once, so eight single-source phases cost more than they hide at world = 8)."""
    max_phases = MAX_PHASES if max_phases is None else max_phases
    peers = [(rank + i) % world for i in range(1, world)]
    groups = [[rank]]
    k = min(len(peers), max(max_phases - 1, 1))
    if k:
        size = -(-len(peers) // k)
        groups += [peers[i:i + size] for i in range(0, len(peers), size)]
    return groups


def split_columns_by_group(rp, c, v, pad: int, groups):
    """One CSR per source group: the entries whose column (padded all-gather layout, owner = column // pad) belongs to a
    source of the group, columns re-based to the group's stage buffer (slot of the source in the group * pad + local
    row).  Entry order inside a row is kept."""
    world = sum(len(g) for g in groups)
    group_of = torch.empty(world, dtype=torch.int64, device=c.device)
    slot_of = torch.empty(world, dtype=torch.int64, device=c.device)
    for gi, g in enumerate(groups):
        for si, src in enumerate(g):
            group_of[src], slot_of[src] = gi, si
    n_rows = rp.numel() - 1
    counts = (rp[1:] - rp[:-1]).long()
    rows = torch.repeat_interleave(torch.arange(n_rows, device=c.device), counts)
    cl = c.long()
    owner = torch.div(cl, pad, rounding_mode="floor")
    new_col = (slot_of[owner] * pad + (cl - owner * pad)).to(torch.int32)
    grp = group_of[owner]
    out = []
    for gi in range(len(groups)):
        sel = grp == gi
        cnt = torch.bincount(rows[sel], minlength=n_rows)
        rp_new = torch.zeros(n_rows + 1, dtype=torch.int64, device=c.device)
        torch.cumsum(cnt, 0, out=rp_new[1:])
        out.append((rp_new.to(torch.int32), new_col[sel].contiguous(), v[sel].contiguous()))
    return out


def _split_columns(rp, c, v, lo_col: int, hi_col: int):
    """(own, rest): the entries whose column lies in [lo_col, hi_col), re-based to 0, and all others (columns unchanged)."""
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


class PeerExchange:
    """Feature-shard exchange of the pipelined row-partitioned SpMM over NVLink peer memory.

    Every rank publishes its [pad, h] shard in a symmetric (peer-mapped) buffer; after one device-side barrier each
    rank PULLS the peers' shards with copy-engine transfers on side streams, one event per source group, in the order
    rank+1, rank+2, ...  so that the SpMM phase of a group starts as soon as its shards have landed while the other
    transfers are still in flight.  Two buffer generations alternate: a rank may publish the next product's shard
    while slower peers still read the previous one (the barrier of product k+1 proves every pull of product k is
    complete before generation k % 2 is written again).  torch.distributed._symmetric_memory only provides the
    allocation, the address exchange and the barrier; no library collective moves data."""

    def __init__(self, pad: int, h: int, rank: int, world: int, group, device, groups):
        import torch.distributed._symmetric_memory as symm_mem
        self.pad, self.h, self.rank, self.world = pad, h, rank, world
        self.group = group if group is not None else dist.group.WORLD
        self.buf = symm_mem.empty((2, pad, h), dtype=torch.float32, device=device)
        self.handle = symm_mem.rendezvous(self.buf, self.group)
        self.groups = groups[1:]                                   # remote source groups (groups[0] is the own shard)
        self.peer = {s: [self.handle.get_buffer(s, (pad, h), torch.float32, g * pad * h) for g in range(2)]
                     for grp in self.groups for s in grp}
        self.stage = [torch.empty((len(grp) * pad, h), dtype=torch.float32, device=device) for grp in self.groups]
        self.streams = [torch.cuda.Stream(device=device) for _ in self.groups]
        self.gen = 0

    def start(self, send: torch.Tensor):
        """send [pad, h] (this rank's padded shard).  Returns [(stage buffer, wait())] per phase, own shard first."""
        main = torch.cuda.current_stream(send.device)
        g = self.gen
        self.gen ^= 1
        self.buf[g].copy_(send)
        self.handle.barrier(channel=0, timeout_ms=20000)     # every rank's generation-g shard is complete (traps, not hangs)
        out = [(send, lambda: None)]
        for gi, grp in enumerate(self.groups):
            st = self.streams[gi]
            st.wait_stream(main)                             # after the barrier, and after the last reader of this stage
            with torch.cuda.stream(st):
                for si, s_ in enumerate(grp):
                    self.stage[gi][si * self.pad:(si + 1) * self.pad].copy_(self.peer[s_][g], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(st)
            out.append((self.stage[gi], (lambda e=ev: main.wait_event(e))))
        return out


class RowPartitionedAdj:
    """This rank's row blocks of A and A^T with columns remapped to the padded all-gather layout.

    overlap=True additionally splits each block by column ownership: the entries that touch this rank's own
    feature rows are multiplied while the all-gather is in flight, the remote ones afterwards (SURVEY.md
    section 8e "start SpMM on locally-owned columns first").  The two partial products are added, so the
    summation order differs from the replicated SpMM (last-bit differences); overlap=False is bit-identical.
    A kernel that gathers straight from peer memory instead is the wrong trade here: peer loads bypass the
    local L2, so every remote entry would pull its 256-byte row over NVLink (12.8 GB per SpMM at N = 2 on the
    stress graph) against 256 MB for the all-gather."""

    def __init__(self, adj: ops.CSRAdj, rank: int, world: int, group=None, overlap: bool = False,
                 pipelined: bool = False):
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
        # pipelined: one sub-block per SOURCE rank (columns re-based to that rank's shard); the product is taken in
        # `world` phases that accumulate into y, phase s as soon as shard s has arrived (PeerExchange)
        self.phases = None
        self.exchange = {}                 # h -> PeerExchange
        self.exchange_override = None      # tests: callable(padded_local) -> [world shards] without peer memory
        if pipelined:
            self.groups = phase_groups(rank, world)
            fwd = split_columns_by_group(*blk, self.pad, self.groups)
            bwd = split_columns_by_group(*blk_t, self.pad, self.groups)
            self.phases = [ops.CSRAdj(self.rows, *f, *b, none, adj.aggr).make_plans() for f, b in zip(fwd, bwd)]
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

    def _pad(self, local: torch.Tensor) -> torch.Tensor:
        if self.rows == self.pad:
            return local.contiguous()
        send = torch.zeros((self.pad, local.shape[1]), dtype=local.dtype, device=local.device)
        send[:self.rows] = local
        return send

    def _spmm_pipelined(self, v_local: torch.Tensor, transposed: bool) -> torch.Tensor:
        h = v_local.shape[1]
        send = self._pad(v_local)
        if self.exchange_override is not None:
            shards = self.exchange_override(send)
            arrivals = [(send, (lambda: None))] + [(torch.cat([shards[s_] for s_ in grp]), (lambda: None))
                                                   for grp in self.groups[1:]]
        elif self.world == 1:
            arrivals = [(send, (lambda: None))]
        else:
            ex = self.exchange.get(h)
            if ex is None:
                ex = self.exchange[h] = PeerExchange(self.pad, h, self.rank, self.world, self.group, v_local.device,
                                                     self.groups)
            arrivals = ex.start(send)
        y = torch.empty((self.rows, h), dtype=torch.float32, device=v_local.device)
        for i, (buf, wait) in enumerate(arrivals):
            wait()
            a = self.phases[i].t() if transposed else self.phases[i]
            ops._run_spmm(a.rowptr, a.col, a.val, a.plan, buf, y, accumulate=i > 0)
        return y

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
        if part.phases is not None:
            return part._spmm_pipelined(x_local, False)
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
        if part.phases is not None:
            return part._spmm_pipelined(gy.contiguous(), True), None
        if part.own is not None:
            return part._spmm_overlapped(gy.contiguous(), True), None
        gy_full = part._gather(gy.contiguous())
        gx = torch.empty((part.rows, gy.shape[1]), dtype=torch.float32, device=gy.device)
        a = part.local
        ops._run_spmm(a.rowptr_t, a.col_t, a.val_t, a.plan_t, gy_full, gx)
        return gx, None


# ---------------------------------------------------------------------------------------------------------
# The whole GLASS model on row shards (stress config): GEMMs, label mix and pooling gathers are
# row-local; per layer one feature all-gather (RowPartitionedAdj.spmm); per GraphNorm one 2C-value fp64 all-reduce of
# the column sums; the pooled subgraph vectors are summed across ranks (a subgraph's nodes live on any rank).
# ---------------------------------------------------------------------------------------------------------
class Comm:
    """The two collectives the partitioned model needs.  Default: torch.distributed on `group`; tests substitute an
    in-process implementation (threads standing in for ranks on one GPU)."""

    def __init__(self, group=None):
        self.group = group
        on = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if on else 1
        self.rank = dist.get_rank(group) if on else 0

    def all_reduce_sum(self, t: torch.Tensor) -> torch.Tensor:
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t


def _stats_total(partial: torch.Tensor, nblk: int, comm: Comm) -> torch.Tensor:
    """Per-block partial column sums of this rank -> global totals as a one-block table [2c, 1] (fp64)."""
    tot = partial[:, :nblk].sum(dim=1, keepdim=True).contiguous()
    return comm.all_reduce_sum(tot)


class _DistGraphNorm(torch.autograd.Function):
    """dropout(act(GraphNorm(x))) with x distributed by rows: local partial sums, one all-reduce, local apply.
    The parameter gradients returned here are this rank's 1/P share of the (already global) sums, so that ONE
    uniform SUM all-reduce over all parameter gradients (PartitionedGLASS.reduce_grads) finishes every gradient."""

    @staticmethod
    def forward(ctx, x, weight, bias, mean_scale, eps, act, p, training, n_total, comm):
        lib = ops._lib.load()
        x, _ = ops._rowmajor(ops._req(x, torch.float32, "x", 2))
        n, c = x.shape
        ld = lib.glass_graphnorm_partials_ld()
        partial = torch.empty((2 * c, ld), dtype=torch.float64, device=x.device)
        nblk = ops.graphnorm_partials(x, partial)
        tot = _stats_total(partial, nblk, comm)
        drop_p, keep, rng, bits = ops._dropout_source(n, c, p, training, x.device)
        stats = torch.empty((6, c), dtype=torch.float32, device=x.device)
        ops._ops.graphnorm_stats_(tot, 1, n_total, weight, bias, mean_scale, float(eps), keep, drop_p, rng, bits, stats)
        out = torch.empty((n, c), dtype=torch.float32, device=x.device)
        ops._ops.graphnorm_apply_(x, stats, act, keep, drop_p, bits, out)
        ctx.save_for_backward(x, weight, mean_scale, stats, keep, bits)
        ctx.cfg = (act, drop_p, n_total, comm)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, weight, mean_scale, stats, keep, bits = ctx.saved_tensors
        act, drop_p, n_total, comm = ctx.cfg
        lib = ops._lib.load()
        dout, _ = ops._rowmajor(dout)
        n, c = x.shape
        partial = torch.empty((2 * c, lib.glass_graphnorm_partials_ld()), dtype=torch.float64, device=x.device)
        nblk = ops.graphnorm_bwd_partials(dout, x, stats, act, keep, drop_p, bits, partial)
        tot = _stats_total(partial, nblk, comm)
        dx = torch.empty((n, c), dtype=torch.float32, device=x.device)
        dw, db, da = torch.empty_like(weight), torch.empty_like(weight), torch.empty_like(weight)
        ops.graphnorm_bwd_finish(tot, n_total, dout, x, weight, mean_scale, stats, act, keep, drop_p, bits, dx, dw, db, da)
        inv = 1.0 / comm.world
        return dx, dw * inv, db * inv, da * inv, None, None, None, None, None, None


class _SumAcrossRanks(torch.autograd.Function):
    """y = sum over ranks of x.  Everything downstream is computed identically on every rank, so the gradient
    of the sum with respect to this rank's term is the incoming gradient itself."""

    @staticmethod
    def forward(ctx, x, comm):
        return comm.all_reduce_sum(x.clone())

    @staticmethod
    def backward(ctx, g):
        return g, None


class PartitionedGLASS:
    """Runs a glass_b200.models.GLASS model (its parameters, replicated on every rank) on ONE rank's row shard.

    forward(h_local, subG_node, z): h_local [rows, H] = this rank's rows of the input embedding (a shard of the table
    -- at stress scale the table itself is sharded), subG_node the GLOBAL padded node ids, z the GLOBAL labels.
    After loss.backward(), reduce_grads() adds the gradients of the row-local operators (GEMM weights) across ranks;
    GraphNorm / head gradients were formed from global sums and are scaled so that the same SUM finishes them."""

    def __init__(self, model, part: RowPartitionedAdj, comm: Comm = None):
        from . import models
        if not isinstance(model, models.GLASS) or model.conv.gns is None:
            raise NotImplementedError("PartitionedGLASS wraps models.GLASS(EmbZGConv(..., gn=True), ...)")
        self.model, self.part = model, part
        self.comm = comm if comm is not None else Comm(part.group)
        self.n = part.n

    def _gn(self, gn, x, act=0, p=0.0):
        return _DistGraphNorm.apply(x, gn.weight, gn.bias, gn.mean_scale, gn.eps, act, p, self.model.training, self.n,
                                    self.comm)

    def _conv(self, conv, h, mask):
        from .models import _act_id
        t0, t1 = conv.trans_fns
        c0, c1 = conv.comb_fns
        x = ops.pair_linear_mix(h, None, t0.weight, t0.bias, t1.weight, t1.bias, mask, conv.z_ratio,
                                _act_id(conv.activation))                                    # impl/models.py:158-162
        x = self.part.spmm(x)                                                                # :164 (all-gather inside)
        x = self._gn(conv.gn, x, p=conv.dropout)                                             # :165-166
        return ops.pair_linear_mix(x, h, c0.weight, c0.bias, c1.weight, c1.bias, mask, conv.z_ratio, 0)   # :167-173

    def node_emb(self, h_local, z=None):
        from .models import _act_id
        net, part = self.model.conv, self.part
        dev = h_local.device
        if z is None:
            mask = torch.ones(part.rows, dtype=torch.uint8, device=dev)
        else:
            mask = ops.label_mask(z)[part.lo:part.hi].contiguous()
        act = _act_id(net.activation)
        h = self._gn(net.emb_gn, h_local, p=net.dropout)                                     # :249-251
        xs = []
        for layer, conv in enumerate(net.convs):
            h = self._conv(conv, h, mask)
            xs.append(h)
            if layer < len(net.convs) - 1:
                h = self._gn(net.gns[layer], h, act=act, p=net.dropout)                      # :257-259
        last = net.gns[-1]
        if net.jk and len(xs) > 1:                                                           # :263-267, per column block
            outs, off = [], 0
            for t in xs:
                w = t.shape[1]
                outs.append(_DistGraphNorm.apply(t, last.weight[off:off + w], last.bias[off:off + w],
                                                 last.mean_scale[off:off + w], last.eps, 0, 0.0, False, self.n,
                                                 self.comm))
                off += w
            return torch.cat(outs, dim=1)
        return self._gn(last, xs[-1])

    def pool(self, emb_local, subG_node, pool):
        """GLASS.Pool over nodes that live on any rank: local segment sums of the owned nodes, summed across ranks."""
        mode = pool.padded_mode()
        if mode not in ("sum", "mean", "size"):
            raise NotImplementedError(f"partitioned pooling supports sum / mean / size, not {mode}")
        part = self.part
        own = (subG_node >= part.lo) & (subG_node < part.hi)
        local = torch.where(own, subG_node - part.lo, torch.full_like(subG_node, -1))
        s = _SumAcrossRanks.apply(ops.segment_pool(emb_local, local, "sum"), self.comm)
        cnt = (subG_node >= 0).sum(dim=1, keepdim=True).to(s.dtype)
        if mode == "mean":
            return s / cnt.clamp(min=1)
        if mode == "size":
            return s * torch.where(cnt > 0, cnt.pow(-0.5), torch.zeros_like(cnt))
        return s

    def forward(self, h_local, subG_node, z=None, id=0):
        emb = self.node_emb(h_local, z)
        return self.model.preds[id](self.pool(emb, subG_node, self.model.pools[id]))

    __call__ = forward

    def reduce_grads(self):
        """SUM the parameter gradients across ranks (call after backward)."""
        inv = 1.0 / self.comm.world
        head = {id(p) for p in self.model.preds.parameters()}
        for p in self.model.parameters():
            if p.grad is None:
                continue
            if id(p) in head:          # computed identically on every rank from the summed pooled vectors
                p.grad.mul_(inv)
            self.comm.all_reduce_sum(p.grad)
