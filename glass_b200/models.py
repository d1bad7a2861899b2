"""Drop-in modules for the GLASS hot path, same API as the reference's impl/models.py.

buildAdj (:83), GLASSConv (:114), EmbZGConv (:177), PoolModule / AddPool / MaxPool / MeanPool /
SizePool (:275-319) and GLASS (:322) keep their constructor signatures, forward signatures, parameter
names (identical ``state_dict`` keys) and parameter creation order (identical seeded initialisation),
but every tensor op runs in the hand-written sm_100a kernels of libglass_b200.so via glass_b200.ops.
There is no torch.sparse / PyG / CPU fallback: CPU tensors raise.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import ops
from ._lib import ACT_ELU, ACT_NONE, ACT_RELU
from .utils import pad2batch


def _act_id(activation) -> int:
    """Map the activation *module* the reference passes (GLASSTest.py:143 nn.ELU(inplace=True); default
    nn.ReLU, impl/models.py:125, 192) to the fused-kernel enum.  Anything else is not accelerated."""
    if activation is None or isinstance(activation, nn.Identity):
        return ACT_NONE
    if isinstance(activation, nn.ReLU):
        return ACT_RELU
    if isinstance(activation, nn.ELU) and activation.alpha == 1.0:
        return ACT_ELU
    raise NotImplementedError(f"activation {activation!r} is not supported by the fused GLASS kernels")


class GraphNorm(nn.Module):
    """Parameter-compatible stand-in for PyG 1.7.2 ``GraphNorm(in_channels)`` (weight, bias, mean_scale
    registered in that order, ones/zeros/ones).  Only whole-graph normalisation (batch=None), which is
    the only way the GLASS path calls it (impl/models.py:165, 249, 257, 266)."""

    def __init__(self, in_channels: int, eps: float = 1e-5):
        super().__init__()
        self.in_channels = in_channels
        self.eps = eps
        self.weight = nn.Parameter(torch.empty(in_channels))
        self.bias = nn.Parameter(torch.empty(in_channels))
        self.mean_scale = nn.Parameter(torch.empty(in_channels))
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.ones_(self.weight)
        nn.init.zeros_(self.bias)
        nn.init.ones_(self.mean_scale)

    def forward(self, x, batch=None, act: int = ACT_NONE, p: float = 0.0, training: bool = False):
        if batch is not None:
            raise NotImplementedError("only whole-graph GraphNorm (batch=None) is on the GLASS path")
        return ops.graph_norm(x, self.weight, self.bias, self.mean_scale, self.eps, act, p, training)


_adj_cache = {}
_fuse_norm_pool = os.environ.get("GLASS_B200_NORM_POOL", "1") != "0"     # last GraphNorm + pooling as one operator


def buildAdj(edge_index, edge_weight, n_node: int, aggr: str):
    """impl/models.py:83-111.  Returns a CSRAdj (CSR + transposed CSR, normalised for ``aggr``) that
    supports ``adj @ x``.  Identical (edge_index, edge_weight, n_node, aggr) inputs share one CSR so
    the L layers of a model do not each rebuild the same matrix."""
    if aggr not in ("mean", "sum", "gcn"):
        raise NotImplementedError
    key = (edge_index.data_ptr(), edge_weight.data_ptr(), tuple(edge_index.shape), int(n_node), aggr,
           edge_index._version, edge_weight._version, str(edge_index.device))
    hit = _adj_cache.get(key)
    if hit is not None and hit[0]() is edge_index and hit[1]() is edge_weight:
        return hit[2]
    adj = ops.build_csr(edge_index, edge_weight, n_node, aggr)
    import weakref
    if len(_adj_cache) > 16:
        _adj_cache.clear()
    _adj_cache[key] = (weakref.ref(edge_index), weakref.ref(edge_weight), adj)
    return adj


def _mask_u8(mask: torch.Tensor) -> torch.Tensor:
    """bool [N,1] (reference convention, impl/models.py:242-246) or uint8 [N] -> uint8 [N] without a copy."""
    m = mask.reshape(-1)
    if m.dtype == torch.bool:
        m = m.view(torch.uint8)
    return m


class GLASSConv(nn.Module):
    """impl/models.py:114-174: label-mixed transform -> adj @ x -> GraphNorm -> dropout -> [x | x_] ->
    label-mixed combine.  Four kernels: pair GEMM (+act+mix), SpMM, GraphNorm(+dropout), pair GEMM over
    the virtual concat (+mix)."""

    def __init__(self, in_channels: int, out_channels: int, activation=nn.ReLU(inplace=True), aggr="mean",
                 z_ratio=0.8, dropout=0.2):
        super().__init__()
        # creation order fixes both the state_dict layout and the seeded initial values: W0, W1 of the
        # transform, W0, W1 of the combine (input = [aggregated | x_]), then the norm
        wide = in_channels + out_channels
        self.trans_fns = nn.ModuleList(nn.Linear(in_channels, out_channels) for _ in range(2))
        self.comb_fns = nn.ModuleList(nn.Linear(wide, out_channels) for _ in range(2))
        self.gn = GraphNorm(out_channels)
        self.activation, self.aggr, self.z_ratio, self.dropout = activation, aggr, z_ratio, dropout
        self.adj = None                     # CSRAdj, built on the first forward and kept (impl/models.py:154-156)
        self.reset_parameters()

    def reset_parameters(self):
        for module in (*self.trans_fns, *self.comb_fns, self.gn):
            module.reset_parameters()

    def forward(self, x_, edge_index, edge_weight, mask):
        if self.adj is None:  # cached on first use, like impl/models.py:154-156
            self.adj = buildAdj(edge_index, edge_weight, x_.shape[0], self.aggr)
        m = _mask_u8(mask)
        t0, t1 = self.trans_fns
        c0, c1 = self.comb_fns
        gn = self.gn
        if x_.is_cuda and x_.dim() == 2 and ops.conv_fusable(x_.shape[1], t0.weight.shape[0]):
            # the whole layer as one fused chain (tcgen05 shapes): ops._GlassConv
            return ops.glass_conv(x_, self.adj, (t0.weight, t0.bias, t1.weight, t1.bias),
                                  (gn.weight, gn.bias, gn.mean_scale, gn.eps),
                                  (c0.weight, c0.bias, c1.weight, c1.bias), m, self.z_ratio,
                                  _act_id(self.activation), self.dropout, self.training)
        x = ops.pair_linear_mix(x_, None, t0.weight, t0.bias, t1.weight, t1.bias, m, self.z_ratio,
                                _act_id(self.activation))                                   # :158-162
        # :164-166, one fused op
        x = ops.spmm_graph_norm(self.adj, x, gn.weight, gn.bias, gn.mean_scale, gn.eps, ACT_NONE, self.dropout,
                                self.training)
        return ops.pair_linear_mix(x, x_, c0.weight, c0.bias, c1.weight, c1.bias, m, self.z_ratio,
                                   ACT_NONE)                                                # :167-173


    # --- multi-label-batch evaluation ---------------------------------------
    # For FIXED weights, the label-mixed features of impl/models.py:161 differ between two label batches only on
    # the labelled rows:  x_b = U + [labelled] * delta  with  U = z*p0 + (1-z)*p1,  delta = (2z-1)*(p1 - p0),
    # so  adj @ x_b = adj @ U + adj[:, labelled] @ delta[labelled]:  one dense SpMM per evaluation epoch, and per
    # batch a correction over the ~1-2 % labelled columns (ops.spmm_delta).  Exact up to fp32 re-association.
    @torch.no_grad()
    def shared_base(self, x_, edge_index, edge_weight):
        """Label-independent part of this layer for the current weights: (x_, adj @ U, delta)."""
        if self.adj is None:
            self.adj = buildAdj(edge_index, edge_weight, x_.shape[0], self.aggr)
        t0, t1 = self.trans_fns
        z = float(self.z_ratio)
        n, h = x_.shape[0], t0.weight.shape[0]
        if h % 4 or h > 128:
            raise NotImplementedError(f"shared-base evaluation needs a hidden width that is a multiple of 4 (<= 128), got {h}")
        none = torch.zeros(n, dtype=torch.uint8, device=x_.device)                 # nobody labelled -> out = U
        u = torch.empty((n, h), dtype=torch.float32, device=x_.device)
        acts = torch.empty((n, 2 * h), dtype=torch.float32, device=x_.device)      # p0 | p1
        ops.pair_linear_mix_into(x_, None, t0.weight, t0.bias, t1.weight, t1.bias, none, z, _act_id(self.activation),
                                 u, acts)
        delta = (2.0 * z - 1.0) * (acts[:, h:] - acts[:, :h])
        return x_, ops.spmm(self.adj, u), delta.contiguous()

    @torch.no_grad()
    def forward_from_base(self, base, mask):
        """Evaluation-mode forward of this layer for one label batch given shared_base()."""
        x_, y_u, delta = base
        m = _mask_u8(mask)
        gn = self.gn
        c0, c1 = self.comb_fns
        return ops.glass_conv_from_base(self.adj, x_, y_u, delta, m, (gn.weight, gn.bias, gn.mean_scale, gn.eps),
                                        (c0.weight, c0.bias, c1.weight, c1.bias), self.z_ratio)


class EmbZGConv(nn.Module):
    """impl/models.py:177-272: embedding -> GraphNorm -> dropout -> L x GLASSConv (GraphNorm / activation /
    dropout between layers) -> optional JK concat -> final GraphNorm."""

    def __init__(self, hidden_channels, output_channels, num_layers, max_deg, dropout=0, activation=nn.ReLU(),
                 conv=GLASSConv, gn=True, jk=False, **kwargs):
        super().__init__()
        out_widths = [hidden_channels] * (num_layers - 1) + [output_channels]       # per conv layer
        self.input_emb = nn.Embedding(max_deg + 1, hidden_channels, scale_grad_by_freq=False)
        self.emb_gn = GraphNorm(hidden_channels)
        self.convs = nn.ModuleList(conv(in_channels=hidden_channels, out_channels=w, activation=activation, **kwargs)
                                   for w in out_widths)
        self.jk, self.activation, self.dropout = jk, activation, dropout
        if gn:   # one norm after every layer; the last one spans the JK concat when jk is set
            norm_widths = out_widths[:-1] + [sum(out_widths) if jk else output_channels]
            self.gns = nn.ModuleList(GraphNorm(w) for w in norm_widths)
        else:
            self.gns = None
        self.reset_parameters()

    def reset_parameters(self):
        for module in (self.input_emb, self.emb_gn, *self.convs, *(self.gns or ())):
            module.reset_parameters()

    def _identity_lookup(self, ids: torch.Tensor) -> bool:
        """--use_nodeid: ids == arange(N) and the table has N rows, so the lookup (impl/models.py:248) is the
        identity.  Checked once per id tensor (one host sync), then cached."""
        table = self.input_emb.weight
        if ids.numel() != table.shape[0]:
            return False
        key = (ids.data_ptr(), ids._version, ids.numel(), str(ids.device))
        hit = getattr(self, "_id_cache", None)
        if hit is None or hit[0] != key:
            same = bool(torch.equal(ids, torch.arange(ids.numel(), device=ids.device, dtype=ids.dtype)))
            self._id_cache = hit = (key, same)
        return hit[1]

    def _mask(self, n, z, device):
        if z is None:  # every node takes the "labelled" branch, impl/models.py:242-244
            return torch.ones(n, dtype=torch.uint8, device=device)
        return ops.label_mask(z)                                                            # :246

    def _input(self, x):
        """:248-251: embedding lookup -> emb_gn -> dropout."""
        if self.gns is None:
            raise NotImplementedError("gn=False is outside the accelerated GLASS path (GLASSTest.py:150 uses gn=True)")
        n = x.shape[0]
        ids = x.reshape(-1)
        if ids.numel() == n and self._identity_lookup(ids):
            h = self.input_emb.weight                                                       # :248, identity gather
        else:
            h = ops.embedding(ids, self.input_emb.weight).reshape(n, -1)                    # :248
        return self.emb_gn(h, p=self.dropout, training=self.training)                       # :249-251

    def _layers(self, h, edge_index, edge_weight, mask, first=None, pool_to=None):
        """:253-272 from the output of _input on; `first` replaces the call of convs[0] (shared-base evaluation).
        pool_to = (subG_node, mode): return the POOLED output of the last GraphNorm instead of the [N, D] embedding --
        that norm feeds nothing but GLASS.Pool (:266 / :272 -> :346-350), so it is applied to the gathered rows only
        (ops.graph_norm_pool)."""
        act = _act_id(self.activation)
        xs = []
        for layer, conv in enumerate(self.convs):
            h = first(mask) if (layer == 0 and first is not None) else conv(h, edge_index, edge_weight, mask)
            xs.append(h)
            if layer < len(self.convs) - 1:                                                 # :253-259
                h = self.gns[layer](h, act=act, p=self.dropout, training=self.training)
        last = self.gns[-1]
        if self.jk and len(xs) > 1:                                                         # :263-267
            if pool_to is not None:
                return ops.graph_norm_pool_cat(xs, last.weight, last.bias, last.mean_scale, last.eps, *pool_to)
            return ops.graph_norm_cat(xs, last.weight, last.bias, last.mean_scale, last.eps)
        if pool_to is not None:
            return ops.graph_norm_pool(xs[-1], last.weight, last.bias, last.mean_scale, last.eps, *pool_to)
        return last(xs[-1])                                                                 # :268-272

    def forward(self, x, edge_index, edge_weight, z=None):
        h = self._input(x)
        return self._layers(h, edge_index, edge_weight, self._mask(x.shape[0], z, x.device))

    def forward_pooled(self, x, edge_index, edge_weight, z, subG_node, mode: str):
        """Pool(forward(x, ...), subG_node) for the padded sum / mean / size pools without materialising the output of
        the last GraphNorm (see _layers)."""
        h = self._input(x)
        return self._layers(h, edge_index, edge_weight, self._mask(x.shape[0], z, x.device), pool_to=(subG_node, mode))

    @torch.no_grad()
    def shared_base(self, x, edge_index, edge_weight):
        """Everything of an evaluation pass that does not depend on the label batch (fixed weights): the normalised
        input embedding and the first layer's adj @ U / delta (GLASSConv.shared_base)."""
        if self.training:
            raise RuntimeError("shared_base is an evaluation-mode fast path (dropout must be off)")
        return self.convs[0].shared_base(self._input(x), edge_index, edge_weight)

    @torch.no_grad()
    def forward_from_base(self, base, edge_index, edge_weight, z=None, pool_to=None):
        n = base[0].shape[0]
        mask = self._mask(n, z, base[0].device)
        return self._layers(None, edge_index, edge_weight, mask,
                            first=lambda m: self.convs[0].forward_from_base(base, m), pool_to=pool_to)


# --- pooling --------------------------------------------------------------------------------------
def global_add_pool(x, batch, size=None):
    return ops.segment_pool_batch(x, batch, "sum", size)


def global_mean_pool(x, batch, size=None):
    return ops.segment_pool_batch(x, batch, "mean", size)


def global_max_pool(x, batch, size=None):
    return ops.segment_pool_batch(x, batch, "max", size)


_POOL_MODE = {global_add_pool: "sum", global_mean_pool: "mean", global_max_pool: "max"}


class PoolModule(nn.Module):
    """impl/models.py:275-292.  ``forward(x, batch)`` pools rows that were already gathered;
    ``pool_padded(emb, subG_node)`` is the fused path GLASS.Pool takes (no pad2batch, no gather)."""

    def __init__(self, pool_fn, trans_fn=None):
        super().__init__()
        self.pool_fn = pool_fn
        self.trans_fn = trans_fn

    def padded_mode(self):
        return _POOL_MODE.get(self.pool_fn) if self.trans_fn is None else None

    def forward(self, x, batch):
        if self.trans_fn is not None:
            x = self.trans_fn(x)
        return self.pool_fn(x, batch)

    def pool_padded(self, emb, subG_node):
        return ops.segment_pool(emb, subG_node, self.padded_mode())


def _pool_class(name: str, pool_fn):
    """AddPool / MaxPool / MeanPool of impl/models.py:295-307: PoolModule with the pooling function bound."""
    def __init__(self, trans_fn=None):
        PoolModule.__init__(self, pool_fn, trans_fn)
    return type(name, (PoolModule,), {"__init__": __init__, "__module__": __name__})


AddPool = _pool_class("AddPool", global_add_pool)
MaxPool = _pool_class("MaxPool", global_max_pool)
MeanPool = _pool_class("MeanPool", global_mean_pool)


class SizePool(AddPool):
    """Sum of rows scaled by (segment size)^-1/2 (GraphSizeNorm then add, impl/models.py:310-319)."""

    def __init__(self, trans_fn=None):
        super().__init__(trans_fn)

    def padded_mode(self):
        return "size" if self.trans_fn is None else None

    def forward(self, x, batch):
        if x is not None and self.trans_fn is not None:
            x = self.trans_fn(x)
        return ops.segment_pool_batch(x, batch, "size")


class _SubgraphModel(nn.Module):
    """Shared plumbing of GLASS and EdgeGNN: node embeddings -> pooled subgraph vectors -> prediction head `id`."""

    def __init__(self, conv, preds: nn.ModuleList, pools: nn.ModuleList):
        super().__init__()
        self.conv, self.preds, self.pools = conv, preds, pools

    def NodeEmb(self, x, edge_index, edge_weight, z=None):
        """impl/models.py:336-344: one pass per feature copy x[:, c, :] (always one copy on this path), averaged."""
        n, copies, width = x.shape
        embs = [self.conv(x[:, c, :].reshape(n, width), edge_index, edge_weight, z) for c in range(copies)]
        # the mean over a single copy is the identity (bit for bit), so no kernel is spent on it
        return embs[0] if copies == 1 else torch.mean(torch.stack(embs, dim=1), dim=1)

    def forward(self, x, edge_index, edge_weight, subG_node, z=None, id=0):
        pool = self.pools[id]
        mode = pool.padded_mode() if isinstance(pool, PoolModule) else None
        if (mode in ops.NORM_POOL_MODES and x.shape[1] == 1 and hasattr(self.conv, "forward_pooled")
                and type(self).Pool is GLASS.Pool and _fuse_norm_pool):
            # one feature copy (always, on this path) and a padded sum / mean / size pool: the last GraphNorm and the
            # pooling run as one operator (same values as Pool(NodeEmb(...)))
            n, _, width = x.shape
            pooled = self.conv.forward_pooled(x[:, 0, :].reshape(n, width), edge_index, edge_weight, z, subG_node, mode)
        else:
            pooled = self.Pool(self.NodeEmb(x, edge_index, edge_weight, z), subG_node, pool)
        return self.preds[id](pooled)


class GLASS(_SubgraphModel):
    """impl/models.py:322-355."""

    @torch.no_grad()
    def shared_base(self, x, edge_index, edge_weight):
        n, copies, width = x.shape
        if copies != 1:
            raise NotImplementedError("shared-base evaluation handles the single feature copy GLASSTest.py produces")
        return self.conv.shared_base(x[:, 0, :].reshape(n, width), edge_index, edge_weight)

    @torch.no_grad()
    def forward_from_base(self, base, edge_index, edge_weight, subG_node, z=None, id=0):
        """model(x, ei, ew, subG_node, z) in evaluation mode with the label-independent work taken from `base`."""
        pool = self.pools[id]
        mode = pool.padded_mode() if isinstance(pool, PoolModule) else None
        if mode in ops.NORM_POOL_MODES and _fuse_norm_pool:        # last GraphNorm + pooling as one operator
            return self.preds[id](self.conv.forward_from_base(base, edge_index, edge_weight, z, pool_to=(subG_node, mode)))
        emb = self.conv.forward_from_base(base, edge_index, edge_weight, z)
        return self.preds[id](self.Pool(emb, subG_node, pool))

    def Pool(self, emb, subG_node, pool):
        mode = pool.padded_mode() if isinstance(pool, PoolModule) else None
        if mode is not None:
            return ops.segment_pool(emb, subG_node, mode)                                   # fused :347-349
        batch, pos = pad2batch(subG_node)                                                   # :347
        return pool(ops.embedding(pos, emb), batch)                                         # :348-349


# --- plain (unlabeled) GNN used by the reference's link-prediction pre-training (GNNEmb.py) -----------
# the single-weight-set subset of the same kernels.  A single Linear is run
# through the label-mixed pair kernel with both slots bound to the same weights, every row "labelled" and
# z = 1, which selects 1*p1 + 0*p0 == p1 exactly (the weight gradient is dW0 + dW1 = 0 + dW).
def _single_linear(a1, a2, lin: nn.Linear, act: int):
    ones = torch.ones(a1.shape[0], dtype=torch.uint8, device=a1.device)
    return ops.pair_linear_mix(a1, a2, lin.weight, lin.bias, lin.weight, lin.bias, ones, 1.0, act)


class MyGCNConv(nn.Module):
    """impl/models.py:361-397: Linear + activation -> adj @ x -> GraphNorm -> [x | x_] -> Linear."""

    def __init__(self, in_channels: int, out_channels: int, activation=nn.ReLU(inplace=True), aggr="mean"):
        super().__init__()
        self.trans_fn = nn.Linear(in_channels, out_channels)
        self.comb_fn = nn.Linear(in_channels + out_channels, out_channels)
        self.gn = GraphNorm(out_channels)
        self.activation, self.aggr, self.adj = activation, aggr, None

    def reset_parameters(self):
        for module in (self.trans_fn, self.comb_fn, self.gn):
            module.reset_parameters()

    def forward(self, x_, edge_index, edge_weight):
        if self.adj is None:
            self.adj = buildAdj(edge_index, edge_weight, x_.shape[0], self.aggr)
        x = _single_linear(x_, None, self.trans_fn, _act_id(self.activation))              # :386-387
        x = ops.spmm(self.adj, x)                                                          # :388
        x = self.gn(x)                                                                     # :389
        return _single_linear(x, x_, self.comb_fn, ACT_NONE)                               # :390-391


class EmbGConv(nn.Module):
    """impl/models.py:400-476.  The element-wise activation / dropout between layers of this pre-training
    model are plain torch ops (it is not on the GLASS hot path); GEMMs, SpMM and GraphNorm use the kernels."""

    def __init__(self, input_channels: int, hidden_channels: int, output_channels: int, num_layers: int, max_deg: int,
                 dropout=0, activation=nn.ReLU(inplace=True), conv=MyGCNConv, gn=True, jk=False, **kwargs):
        super().__init__()
        # layer widths: input -> hidden -> ... -> hidden -> output (a single layer maps input -> output)
        dims = [input_channels] + [hidden_channels] * (num_layers - 1) + [output_channels]
        self.input_emb = nn.Embedding(max_deg + 1, hidden_channels)
        self.convs = nn.ModuleList(conv(in_channels=a, out_channels=b, **kwargs) for a, b in zip(dims, dims[1:]))
        self.jk, self.activation, self.dropout = jk, activation, dropout
        self.gns = nn.ModuleList(GraphNorm(hidden_channels) for _ in range(num_layers - 1)) if gn else None
        self.reset_parameters()

    def reset_parameters(self):
        for module in (*self.convs, *(self.gns or ())):
            module.reset_parameters()

    def forward(self, x, edge_index, edge_weight, z=None):
        import torch.nn.functional as F
        xs = []
        x = F.dropout(ops.embedding(x.reshape(-1), self.input_emb.weight), p=self.dropout, training=self.training)
        for layer, conv in enumerate(self.convs[:-1]):
            x = conv(x, edge_index, edge_weight)
            if self.gns is not None:
                x = self.gns[layer](x)
            xs.append(x)
            # like the reference, an in-place activation module also rewrites the tensor just appended to xs
            x = F.dropout(self.activation(x), p=self.dropout, training=self.training)
        xs.append(self.convs[-1](x, edge_index, edge_weight))
        return torch.cat(xs, dim=-1) if self.jk else xs[-1]


class EdgeGNN(_SubgraphModel):
    """impl/models.py:479-509: node embeddings -> mean of the two end points of each pair -> preds[id]."""

    def Pool(self, emb, subG_node, pool):
        return ops.segment_pool(emb, subG_node, "mean")      # emb[subG_node].mean(dim=1), impl/models.py:502-504
