"""CPU oracle for the GLASS labeled message-passing hot path.

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module.
glass_b200/ never imports it and has no CPU fallback.

This is a functional (state_dict-in, tensors-out) restatement, in plain PyTorch
CPU fp32 ops, of the algorithm in the reference's impl/models.py and
impl/utils.py plus the PyG 1.7.2 pieces it calls (GraphNorm, GraphSizeNorm,
global_*_pool -- not vendored in /root/reference; formulas restated from the PyG
1.7.2 docs / the GraphNorm paper).  Every function cites the
reference file:line it follows.  It deliberately executes the same ATen op
sequence as the reference (sparse-COO @ dense, `mean[batch]` gathers, two
Linear calls + torch.where) so that timing it on host cores is a faithful CPU
baseline ("port") when /root/reference is not present (it is absent on the GPU
box).  Every function follows its inputs' device, so the same op sequence can also be timed as the "eager
torch.sparse on the GPU" baseline (bench.py --gpu-eager-baseline); the parity tests always run it on the CPU.

Pinning: the reference has no tests or golden vectors for this path
(SURVEY.md section 4).  The oracle is pinned instead against outputs of the
*unmodified reference run in the build container* (tests/golden/make_golden.py
imports /root/reference with oracle/pyg_shim on the path and stores inputs,
state_dicts, outputs and gradients); tests/test_oracle_golden.py replays them.
The PyG stand-in itself is unpinned (no PyG source offline) -- DESIGN.md says so.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# impl/utils.py
# --------------------------------------------------------------------------------------


def pad2batch(pad: torch.Tensor):
    """impl/utils.py:18-29 -- (batch_ids, node_ids) of the entries >= 0, row-major."""
    rows = torch.arange(pad.shape[0], device=pad.device).reshape(-1, 1).expand_as(pad).flatten()
    flat = pad.flatten()
    keep = flat >= 0
    return rows[keep], flat[keep]


def max_zero_one(n_node: int, pos: torch.Tensor) -> torch.Tensor:
    """impl/utils.py:32-45 (MaxZOZ) -- z[n] = 1 iff node n occurs in any row of pos; int64 [N]."""
    z = torch.zeros(n_node, dtype=torch.int64, device=pos.device)
    flat = pos.flatten()
    z[flat[flat >= 0]] = 1
    return z


# --------------------------------------------------------------------------------------
# impl/models.py:83-111  buildAdj
# --------------------------------------------------------------------------------------


def build_adj(edge_index: torch.Tensor, edge_weight: torch.Tensor, n_node: int, aggr: str):
    """impl/models.py:83-111 -- uncoalesced sparse COO, same op sequence as the reference."""
    adj = torch.sparse_coo_tensor(edge_index, edge_weight, size=(n_node, n_node))
    deg = torch.sparse.sum(adj, dim=(1,)).to_dense().flatten()          # :93
    deg[deg < 0.5] += 1.0                                                # :94
    if aggr == "mean":
        dinv = 1.0 / deg                                                 # :96
        val = dinv[edge_index[0]] * edge_weight                          # :98
    elif aggr == "sum":
        val = edge_weight                                                # :102
    elif aggr == "gcn":
        dinv = torch.pow(deg, -0.5)                                      # :105
        val = dinv[edge_index[0]] * edge_weight * dinv[edge_index[1]]    # :107-108
    else:
        raise NotImplementedError                                        # :111
    return torch.sparse_coo_tensor(edge_index, val, size=(n_node, n_node))


def build_csr_numpy(edge_index: np.ndarray, edge_weight: np.ndarray, n_node: int, aggr: str):
    """Bit-exact CSR statement of impl/models.py:83-111 followed by ``.coalesce()``.

    Pure numpy float32/int arithmetic.  Rules (SURVEY.md section 8c "semantics worth pinning"):
      * degree = sum of the RAW entries of a row (duplicates counted), rows with deg < 0.5 get +1;
      * mean: val_e = (1/deg)[row_e] * w_e ; gcn: val_e = ((deg^-1/2)[row_e] * w_e) * (deg^-1/2)[col_e]
        with deg^-1/2 == IEEE 1/sqrt in fp32 (CPU torch.pow(d, -0.5));
      * each raw entry is normalised first, THEN duplicates (same row, col) are summed, in input order;
      * entries sorted by (row, col); gcn adds no self loops.
    Summation order for non-unit weights: sequential over the (row, col)-stable-sorted entries.
    Returns dict(rowptr int32 [N+1], col int32, val float32, deg float32 (after the +1 fix),
    and the transposed CSR rowptr_t/col_t/val_t).
    """
    row = np.asarray(edge_index[0], dtype=np.int64)
    col = np.asarray(edge_index[1], dtype=np.int64)
    w = np.asarray(edge_weight, dtype=np.float32)
    order = np.lexsort((col, row))  # stable: by row, then col, ties keep input order
    row, col, w = row[order], col[order], w[order]
    deg = np.zeros(n_node, dtype=np.float32)
    np.add.at(deg, row, w)  # sequential fp32 accumulation in sorted order
    deg[deg < 0.5] += np.float32(1.0)
    if aggr == "mean":
        dinv = (np.float32(1.0) / deg).astype(np.float32)
        val = (dinv[row] * w).astype(np.float32)
    elif aggr == "sum":
        val = w.copy()
    elif aggr == "gcn":
        dinv = (np.float32(1.0) / np.sqrt(deg, dtype=np.float32)).astype(np.float32)
        val = ((dinv[row] * w).astype(np.float32) * dinv[col]).astype(np.float32)
    else:
        raise NotImplementedError
    key = row * n_node + col
    head = np.ones(key.shape[0], dtype=bool)
    head[1:] = key[1:] != key[:-1]
    seg = np.cumsum(head) - 1
    n_out = int(seg[-1]) + 1 if key.shape[0] else 0
    mval = np.zeros(n_out, dtype=np.float32)
    np.add.at(mval, seg, val)  # sequential fp32 sum of duplicates
    mrow, mcol = row[head], col[head]
    rowptr = np.zeros(n_node + 1, dtype=np.int64)
    np.add.at(rowptr, mrow + 1, 1)
    rowptr = np.cumsum(rowptr)
    # transposed copy: entries sorted by (col, row)
    order_t = np.lexsort((mrow, mcol))
    rowptr_t = np.zeros(n_node + 1, dtype=np.int64)
    np.add.at(rowptr_t, mcol + 1, 1)
    rowptr_t = np.cumsum(rowptr_t)
    return dict(rowptr=rowptr.astype(np.int32), col=mcol.astype(np.int32), val=mval, deg=deg,
                rowptr_t=rowptr_t.astype(np.int32), col_t=mrow[order_t].astype(np.int32),
                val_t=mval[order_t])


# --------------------------------------------------------------------------------------
# PyG 1.7.2 pieces (not vendored in the reference): GraphNorm / GraphSizeNorm / global pools
# --------------------------------------------------------------------------------------


def graph_norm(x, weight, bias, mean_scale, eps: float = 1e-5):
    """PyG GraphNorm.forward with batch=None (call sites impl/models.py:165, 249, 257, 266).

    Same op sequence as PyG: scatter_mean -> mean[batch] gather -> centred var -> affine.
    """
    n = x.shape[0]
    batch = torch.zeros(n, dtype=torch.long, device=x.device)
    mean = torch.zeros(1, x.shape[1], dtype=x.dtype, device=x.device).index_add_(0, batch, x) / n
    out = x - mean[batch] * mean_scale
    var = torch.zeros(1, x.shape[1], dtype=x.dtype, device=x.device).index_add_(0, batch, out * out) / n
    std = (var + eps).sqrt()[batch]
    return weight * out / std + bias


def _segment_count(batch, n_seg, dtype):
    return torch.zeros(n_seg, dtype=dtype, device=batch.device).index_add_(
        0, batch, torch.ones(batch.shape[0], dtype=dtype, device=batch.device))


def pool_nodes(x, batch, kind: str, n_seg: Optional[int] = None):
    """impl/models.py:275-319 -- Add/Mean/Max/Size pooling of gathered rows by segment id."""
    n_seg = int(batch.max()) + 1 if n_seg is None else n_seg
    if kind == "size":  # SizePool: GraphSizeNorm then add (impl/models.py:314-319)
        cnt = _segment_count(batch, n_seg, x.dtype)
        x = x * cnt.pow(-0.5)[batch].view(-1, 1)
        kind = "sum"
    if kind == "sum":
        return torch.zeros(n_seg, x.shape[1], dtype=x.dtype, device=x.device).index_add_(0, batch, x)
    if kind == "mean":
        cnt = _segment_count(batch, n_seg, x.dtype).clamp(min=1)
        return torch.zeros(n_seg, x.shape[1], dtype=x.dtype, device=x.device).index_add_(0, batch, x) / cnt.view(-1, 1)
    if kind == "max":
        idx = batch.view(-1, 1).expand_as(x)
        return torch.zeros(n_seg, x.shape[1], dtype=x.dtype, device=x.device).scatter_reduce(
            0, idx, x, reduce="amax", include_self=False)
    raise NotImplementedError  # GLASSTest.py:171


# --------------------------------------------------------------------------------------
# impl/models.py:114-174  GLASSConv ; :177-272 EmbZGConv ; :322-355 GLASS
# --------------------------------------------------------------------------------------

_ACTS = {"elu": F.elu, "relu": F.relu, "none": lambda t: t}


@dataclass
class GlassConfig:
    """Hyper-parameters that GLASSTest.buildModel (GLASSTest.py:129-175) passes down."""
    hidden_dim: int = 64
    conv_layer: int = 1
    aggr: str = "mean"
    z_ratio: float = 0.8
    dropout: float = 0.0
    pool: str = "sum"
    jk: bool = True
    activation: str = "elu"
    out_dim: int = 1
    gn: bool = True


def _drop(x, p, training, keep):
    """F.dropout (impl/models.py:166, 251, 259) with an explicit keep-mask when one is injected."""
    if not training or p == 0.0:
        return x
    if keep is None:
        return F.dropout(x, p=p, training=True)
    return x * keep.to(x.dtype) / (1.0 - p)


def _mix(x0, x1, mask, z):
    """impl/models.py:161-162 / 172-173."""
    return torch.where(mask, z * x1 + (1 - z) * x0, z * x0 + (1 - z) * x1)


def glass_conv(sd: Dict[str, torch.Tensor], prefix: str, x_, adj, mask, cfg: GlassConfig,
               training: bool, keep=None):
    """impl/models.py:153-174 (GLASSConv.forward); adj is the cached buildAdj result (:154-156)."""
    act = _ACTS[cfg.activation]
    lin = lambda name, t: F.linear(t, sd[f"{prefix}.{name}.weight"], sd[f"{prefix}.{name}.bias"])
    x1 = act(lin("trans_fns.1", x_))                                     # :158
    x0 = act(lin("trans_fns.0", x_))                                     # :159
    x = _mix(x0, x1, mask, cfg.z_ratio)                                  # :161
    x = adj @ x                                                          # :164
    x = graph_norm(x, sd[f"{prefix}.gn.weight"], sd[f"{prefix}.gn.bias"],
                   sd[f"{prefix}.gn.mean_scale"])                        # :165
    x = _drop(x, cfg.dropout, training, keep)                            # :166
    x = torch.cat((x, x_), dim=-1)                                       # :167
    x1 = lin("comb_fns.1", x)                                            # :169
    x0 = lin("comb_fns.0", x)                                            # :170
    return _mix(x0, x1, mask, cfg.z_ratio)                               # :172


def emb_zg_conv(sd, x_ids, adj, z, cfg: GlassConfig, training: bool,
                keeps: Optional[Sequence[torch.Tensor]] = None, prefix: str = "conv"):
    """impl/models.py:240-272 (EmbZGConv.forward).

    keeps: optional injected dropout keep-masks, in consumption order:
      [after emb_gn, (layer l: inside conv, after inter-layer act) for l < L-1 ..., inside last conv].
    """
    keeps = list(keeps) if keeps is not None else None
    nxt = (lambda: keeps.pop(0)) if keeps is not None else (lambda: None)
    n = x_ids.shape[0]
    if z is None:
        mask = torch.ones(n, 1, dtype=torch.bool, device=x_ids.device)   # :242-244
    else:
        mask = (z > 0.5).reshape(-1, 1)                                  # :246
    act = _ACTS[cfg.activation]
    gn = lambda name, t: graph_norm(t, sd[f"{prefix}.{name}.weight"], sd[f"{prefix}.{name}.bias"],
                                    sd[f"{prefix}.{name}.mean_scale"])
    x = F.embedding(x_ids, sd[f"{prefix}.input_emb.weight"]).reshape(n, -1)   # :248
    x = gn("emb_gn", x)                                                  # :249
    x = _drop(x, cfg.dropout, training, nxt())                           # :251
    xs = []
    L = cfg.conv_layer
    for layer in range(L - 1):                                           # :253-259
        x = glass_conv(sd, f"{prefix}.convs.{layer}", x, adj, mask, cfg, training, nxt())
        xs.append(x)
        if cfg.gn:
            x = gn(f"gns.{layer}", x)
        x = act(x)
        x = _drop(x, cfg.dropout, training, nxt())
    x = glass_conv(sd, f"{prefix}.convs.{L - 1}", x, adj, mask, cfg, training, nxt())   # :260
    xs.append(x)
    x = torch.cat(xs, dim=-1) if cfg.jk else xs[-1]                      # :263-269
    if cfg.gn:
        x = gn(f"gns.{L - 1}", x)
    return x


def glass_forward(sd, x, adj, subG_node, z, cfg: GlassConfig, training: bool = False, keeps=None):
    """impl/models.py:352-355 with NodeEmb (:336-344) and Pool (:346-350); head = preds.0 Linear."""
    embs = []
    for c in range(x.shape[1]):                                          # :338
        ids = x[:, c, :].reshape(x.shape[0], x.shape[-1])
        embs.append(emb_zg_conv(sd, ids, adj, z, cfg, training, keeps).unsqueeze(1))
    emb = torch.mean(torch.cat(embs, dim=1), dim=1)                      # :342-343
    batch, pos = pad2batch(subG_node)                                    # :347
    pooled = pool_nodes(emb[pos], batch, cfg.pool)                       # :348-349
    logits = F.linear(pooled, sd["preds.0.weight"], sd["preds.0.bias"])  # :355
    return logits, pooled, emb


def init_state_dict(cfg: GlassConfig, n_emb_rows: int, seed: int = 0, pretrained: Optional[torch.Tensor] = None):
    """Parameters in the reference's creation order and with its initialisers
    (impl/models.py:130-151, 198-229; GLASSTest.py:153-160) so a seed gives identical weights."""
    g = torch.Generator().manual_seed(seed)
    H, L = cfg.hidden_dim, cfg.conv_layer
    sd: Dict[str, torch.Tensor] = {}

    def linear(name, fan_in, fan_out):
        bound = 1.0 / math.sqrt(fan_in)  # kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in))
        sd[f"{name}.weight"] = (torch.rand(fan_out, fan_in, generator=g) * 2 - 1) * bound
        sd[f"{name}.bias"] = (torch.rand(fan_out, generator=g) * 2 - 1) * bound

    def gnorm(name, c):
        sd[f"{name}.weight"] = torch.ones(c)
        sd[f"{name}.bias"] = torch.zeros(c)
        sd[f"{name}.mean_scale"] = torch.ones(c)

    sd["conv.input_emb.weight"] = (pretrained.clone() if pretrained is not None
                                   else torch.randn(n_emb_rows, H, generator=g))
    gnorm("conv.emb_gn", H)
    for l in range(L):
        p = f"conv.convs.{l}"
        linear(f"{p}.trans_fns.0", H, H)
        linear(f"{p}.trans_fns.1", H, H)
        linear(f"{p}.comb_fns.0", 2 * H, H)
        linear(f"{p}.comb_fns.1", 2 * H, H)
        gnorm(f"{p}.gn", H)
    for l in range(L - 1):
        gnorm(f"conv.gns.{l}", H)
    gnorm(f"conv.gns.{L - 1}", H * L if cfg.jk else H)
    linear("preds.0", H * L if cfg.jk else H, cfg.out_dim)
    return sd


def loss_fn_for(out_dim_is_binary: bool):
    """GLASSTest.py:55-71 -- BCEWithLogits on flattened tensors for binary, CrossEntropy otherwise."""
    if out_dim_is_binary:
        return lambda pred, y: F.binary_cross_entropy_with_logits(pred.flatten(), y.flatten())
    return lambda pred, y: F.cross_entropy(pred, y)


@dataclass
class OracleModel:
    """A trainable CPU instance of the oracle (used as the CPU baseline 'port' by bench.py)."""
    cfg: GlassConfig
    sd: Dict[str, torch.Tensor]
    adj: Optional[torch.Tensor] = None
    params: List[torch.Tensor] = field(default_factory=list)

    def __post_init__(self):
        for k in self.sd:
            self.sd[k] = self.sd[k].detach().clone().requires_grad_(True)
        self.params = list(self.sd.values())

    def step(self, optimizer, x, edge_index, edge_weight, pos, y, loss_fn, training=True):
        """One impl/train.py:10-16 iteration (zero_grad / forward / loss / backward / item / step)."""
        if self.adj is None:
            self.adj = build_adj(edge_index, edge_weight, x.shape[0], self.cfg.aggr)
        z = max_zero_one(x.shape[0], pos)                                # SubGDataset.py:92-96
        if training:
            optimizer.zero_grad()
            logits, _, _ = glass_forward(self.sd, x, self.adj, pos, z, self.cfg, training=True)
            loss = loss_fn(logits, y)
            loss.backward()
            out = loss.detach().item()
            optimizer.step()
            return out
        with torch.no_grad():
            logits, _, _ = glass_forward(self.sd, x, self.adj, pos, z, self.cfg, training=False)
        return logits
