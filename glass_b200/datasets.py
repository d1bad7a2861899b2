"""Base-graph container and dataset loading for the GLASS hot path.

Mirrors the *interface* of the reference's datasets.py (BaseGraph :11-101, load_dataset :103-229)
for the parts the hot path consumes: x / edge_index / edge_attr / pos / y / mask, setOneFeature,
setNodeIdFeature, setDegreeFeature, get_split, to().  One-time host preprocessing; not accelerated.

Data sources
  * "density", "cut_ratio", "coreness", "component": the reference's shipped synthetic datasets,
    re-encoded as integer arrays in data/<name>.npz by scripts/convert_shipped_datasets.py
    (datasets.py:105-125 reads the same content from a networkx pickle).
  * "ppi_bp_shaped", "em_user_shaped", "stress[_small]": seeded synthetic graphs of the shapes
    BASELINE.json names (the real ppi_bp / em_user data and embeddings are not shipped).
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np
import torch

_DATA_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data")
SHIPPED = ("density", "cut_ratio", "coreness", "component")


def coalesce_undirected(edge: torch.Tensor, weight: torch.Tensor, n_node: int):
    """Symmetrise then sort by (row, col) and merge duplicates by summing weights.

    Same result as PyG ``to_undirected(edge_index, edge_attr)`` used at datasets.py:68-71.
    """
    row = torch.cat((edge[0], edge[1]))
    col = torch.cat((edge[1], edge[0]))
    w = torch.cat((weight, weight))
    key, inv = torch.unique(row * n_node + col, sorted=True, return_inverse=True)
    out_w = torch.zeros(key.numel(), dtype=weight.dtype, device=weight.device).index_add_(0, inv, w)
    return torch.stack((torch.div(key, n_node, rounding_mode="floor"), key % n_node)), out_w


def _is_undirected(edge: torch.Tensor, n_node: int) -> bool:
    key = torch.unique(edge[0] * n_node + edge[1])
    key_t = torch.unique(edge[1] * n_node + edge[0])
    return key.numel() == edge.shape[1] and torch.equal(key, key_t)


class BaseGraph:
    """x [N,1,F] int64 ids, edge_index [2,nnz] int64, edge_attr [nnz] fp32, pos [S,Lmax] (-1 pad),
    y [S], mask [S] in {0: train, 1: valid, 2: test} -- the layout of datasets.py:14-21."""

    def __init__(self, x, edge_index, edge_weight, subG_node, subG_label, mask):
        self.x = x
        self.edge_index = edge_index
        self.edge_attr = edge_weight
        self.pos = subG_node
        self.y = subG_label
        self.mask = mask
        self.to_undirected()

    @property
    def num_nodes(self) -> int:
        return self.x.shape[0]

    def to_undirected(self):
        """datasets.py:68-71.  Edge lists that already live on a GPU are symmetrised / sorted / merged by the CUDA
        kernels (ops.to_undirected); host tensors take the torch path."""
        n = self.x.shape[0]
        if self.edge_index.is_cuda:
            from . import ops
            self.edge_index, self.edge_attr = ops.to_undirected(self.edge_index, self.edge_attr.to(torch.float32), n)
        elif not _is_undirected(self.edge_index, n):
            self.edge_index, self.edge_attr = coalesce_undirected(self.edge_index, self.edge_attr, n)

    def _degree(self):
        deg = torch.zeros(self.x.shape[0], dtype=self.edge_attr.dtype)
        return deg.index_add_(0, self.edge_index[0].cpu(), self.edge_attr.cpu()).to(torch.int64)

    def setDegreeFeature(self, mod: int = 1):
        deg = torch.div(self._degree(), mod, rounding_mode="floor")
        self.x = torch.unique(deg, return_inverse=True)[1].reshape(self.x.shape[0], 1, -1)

    def setOneFeature(self):
        self.x = torch.ones((self.x.shape[0], 1, 1), dtype=torch.int64)

    def setNodeIdFeature(self):
        self.x = torch.arange(self.x.shape[0], dtype=torch.int64).reshape(self.x.shape[0], 1, -1)

    def get_split(self, split: str):
        sel = self.mask == {"train": 0, "valid": 1, "test": 2}[split]
        return self.x, self.edge_index, self.edge_attr, self.pos[sel], self.y[sel]

    def to(self, device):
        for name in ("x", "edge_index", "edge_attr", "pos", "y", "mask"):
            setattr(self, name, getattr(self, name).to(device))
        return self


def load_edges(name: str) -> Tuple[torch.Tensor, torch.Tensor, int]:
    """Undirected, (row, col)-sorted edge_index / unit weights / node count of a shipped dataset."""
    d = np.load(os.path.join(_DATA_DIR, f"{name}.npz"))
    n = int(d["n_node"])
    e = torch.from_numpy(d["edge"].astype(np.int64))
    ei, ew = coalesce_undirected(e, torch.ones(e.shape[1]), n)
    return ei, ew, n


def _pad_rows(rows, fill=-1) -> torch.Tensor:
    lmax = max(len(r) for r in rows)
    out = torch.full((len(rows), lmax), fill, dtype=torch.int64)
    for i, r in enumerate(rows):
        out[i, :len(r)] = torch.as_tensor(r, dtype=torch.int64)
    return out


def uniform_edges(n: int, n_und: int, seed: int) -> torch.Tensor:
    """n_und distinct undirected edges without self loops, sampled uniformly (numpy PCG64, seeded)."""
    g = np.random.default_rng(seed)
    got = np.zeros(0, dtype=np.int64)
    while got.shape[0] < n_und:
        m = int((n_und - got.shape[0]) * 1.1) + 1024
        a, b = g.integers(0, n, m), g.integers(0, n, m)
        keep = a != b
        lo, hi = np.minimum(a, b)[keep], np.maximum(a, b)[keep]
        got = np.unique(np.concatenate((got, lo * n + hi)))
    got = g.permutation(got)[:n_und]
    return torch.from_numpy(np.stack((got // n, got % n)))


def powerlaw_edges(n: int, n_und: int, seed: int, alpha: float = 2.1) -> torch.Tensor:
    """Chung-Lu style skewed graph: endpoints drawn with probability ~ rank^(-1/(alpha-1))."""
    g = np.random.default_rng(seed)
    wgt = np.arange(1, n + 1, dtype=np.float64) ** (-1.0 / (alpha - 1.0))
    cdf = np.cumsum(wgt)
    cdf /= cdf[-1]
    perm = g.permutation(n)  # hubs scattered over the id range
    got = np.zeros(0, dtype=np.int64)
    while got.shape[0] < n_und:
        m = int((n_und - got.shape[0]) * 1.3) + 1024
        a = perm[np.searchsorted(cdf, g.random(m))]
        b = perm[np.searchsorted(cdf, g.random(m))]
        keep = a != b
        lo, hi = np.minimum(a, b)[keep], np.maximum(a, b)[keep]
        got = np.unique(np.concatenate((got, lo * n + hi)))
    got = g.permutation(got)[:n_und]
    return torch.from_numpy(np.stack((got // n, got % n)))


def powerlaw_edges_torch(n: int, n_und: int, seed: int, device, alpha: float = 2.1) -> torch.Tensor:
    """Same Chung-Lu construction as powerlaw_edges, vectorised in torch so that the 100M-edge stress graph
    is generated on the GPU in about a second (numpy needs minutes).  Different RNG stream than the numpy
    generator: the two produce different (equally distributed) graphs."""
    g = torch.Generator(device=device).manual_seed(seed)
    wgt = torch.arange(1, n + 1, dtype=torch.float64, device=device) ** (-1.0 / (alpha - 1.0))
    cdf = torch.cumsum(wgt, 0)
    cdf /= cdf[-1].clone()
    perm = torch.randperm(n, generator=g, device=device)
    got = torch.zeros(0, dtype=torch.int64, device=device)
    while got.numel() < n_und:
        m = int((n_und - got.numel()) * 1.3) + 1024
        a = perm[torch.searchsorted(cdf, torch.rand(m, generator=g, device=device, dtype=torch.float64)).clamp_(max=n - 1)]
        b = perm[torch.searchsorted(cdf, torch.rand(m, generator=g, device=device, dtype=torch.float64)).clamp_(max=n - 1)]
        keep = a != b
        lo, hi = torch.minimum(a, b)[keep], torch.maximum(a, b)[keep]
        got = torch.unique(torch.cat((got, lo * n + hi)))
        del a, b, keep, lo, hi
    got = got[torch.randperm(got.numel(), generator=g, device=device)[:n_und]]
    return torch.stack((torch.div(got, n, rounding_mode="floor"), got % n))


def _random_subgraphs(n, count, mean_len, std_len, min_len, seed):
    g = np.random.default_rng(seed)
    rows = []
    for _ in range(count):
        if std_len is None:
            k = int(g.poisson(mean_len))
        else:
            k = int(round(g.normal(mean_len, std_len)))
        k = min(max(k, min_len), n)
        rows.append(g.choice(n, size=k, replace=False))
    return rows


_SHAPES = {
    # name: nodes, undirected edges, subgraphs, mean/std/min size, classes (2 => binary float labels)
    "ppi_bp_shaped": dict(n=17080, e=316951, s=1591, mean=10, std=None, lmin=2, classes=6),
    "em_user_shaped": dict(n=57333, e=4573417, s=324, mean=155, std=100, lmin=2, classes=2),
    "em_user_shaped_powerlaw": dict(n=57333, e=4573417, s=324, mean=155, std=100, lmin=2, classes=2, gen="powerlaw"),
    "stress": dict(n=2_000_000, e=100_000_000, s=512, mean=64, std=None, lmin=2, classes=2, gen="powerlaw"),
    "stress_small": dict(n=200_000, e=5_000_000, s=256, mean=64, std=None, lmin=2, classes=2, gen="powerlaw"),
}


def synthetic_graph(name: str, seed: int = 0, device=None) -> "BaseGraph":
    """Seeded synthetic stand-ins for the datasets that are not shipped.
    With a CUDA `device`, power-law graphs are generated on the GPU (stress graph: seconds instead of minutes)."""
    s = _SHAPES[name]
    if device is not None and s.get("gen") == "powerlaw":
        e = powerlaw_edges_torch(s["n"], s["e"], seed, device)
    else:
        gen = powerlaw_edges if s.get("gen") == "powerlaw" else uniform_edges
        e = gen(s["n"], s["e"], seed)
    rows = _random_subgraphs(s["n"], s["s"], s["mean"], s["std"], s["lmin"], seed + 1)
    g = np.random.default_rng(seed + 2)
    label = torch.from_numpy(g.integers(0, s["classes"], s["s"]))
    n_trn = int(0.8 * s["s"])
    n_val = int(0.1 * s["s"])
    mask = torch.cat((torch.zeros(n_trn, dtype=torch.int64), torch.ones(n_val, dtype=torch.int64),
                      2 * torch.ones(s["s"] - n_trn - n_val, dtype=torch.int64)))
    return BaseGraph(torch.empty((s["n"], 1, 0)), e, torch.ones(e.shape[1], device=e.device), _pad_rows(rows),
                     label.to(torch.float) if s["classes"] == 2 else label, mask)


def synthetic_embedding(n: int, dim: int, seed: int = 0, std: float = 2.0) -> torch.Tensor:
    """Stand-in for the missing Emb/<dataset>_64.pt (the shipped hpo tables have std 1.8-2.2)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, dim, generator=g) * std


def load_dataset(name: str, seed: Optional[int] = None, device=None) -> "BaseGraph":
    """datasets.py:103-126 for the shipped synthetic sets; seeded generators for the *_shaped ones.

    For shipped sets the split permutation is drawn from torch's global RNG exactly as the
    reference does (datasets.py:118-123), so ``--use_seed`` reproduces the reference's splits.
    """
    if name in SHIPPED:
        d = np.load(os.path.join(_DATA_DIR, f"{name}.npz"))
        n = int(d["n_node"])
        pad = torch.from_numpy(d["subG_pad"].astype(np.int64))
        label = torch.from_numpy(d["label"].astype(np.int64))
        cnt = pad.shape[0]
        mask = torch.cat((torch.zeros(cnt - cnt // 2, dtype=torch.int64),
                          torch.ones(cnt // 4, dtype=torch.int64),
                          2 * torch.ones(cnt // 2 - cnt // 4, dtype=torch.int64)))
        mask = mask[torch.randperm(mask.shape[0])]
        e = torch.from_numpy(d["edge"].astype(np.int64))
        return BaseGraph(torch.empty((n, 1, 0)), e, torch.ones(e.shape[1]), pad, label, mask)
    if name in _SHAPES:
        return synthetic_graph(name, 0 if seed is None else seed, device)
    if name in REAL:
        return load_real_dataset(name)
    raise NotImplementedError(name)


REAL = ("ppi_bp", "hpo_metab", "hpo_neuro", "em_user")


def real_dataset_dir(name: str) -> str:
    """./dataset/<name> like the reference (datasets.py:180-222, CWD-relative); GLASS_DATASET_DIR overrides the root."""
    return os.path.join(os.environ.get("GLASS_DATASET_DIR", "./dataset"), name)


def parse_subgraph_file(path: str):
    """SubGNN `subgraphs.pth` text format (datasets.py:131-174): one subgraph per line,
    `n1-n2-...<TAB>label[-label...]<TAB>train|val|test`.  Label names get ids in order of first appearance;
    more than one label on any line makes the task multi-label.  Returns
    ({split: (node lists, label-id lists)}, multilabel)."""
    label_id = {}
    splits = {"train": ([], []), "val": ([], []), "test": ([], [])}
    multilabel = False
    with open(path) as f:
        for line in f:
            fields = line.split("\t")
            nodes = [int(tok) for tok in fields[0].split("-") if tok != ""]
            if not nodes:
                continue
            names = fields[1].split("-")
            multilabel = multilabel or len(names) > 1
            for nm in names:
                label_id.setdefault(nm, len(label_id))
            split = fields[2].strip()
            if split in splits:
                splits[split][0].append(nodes)
                splits[split][1].append([label_id[nm] for nm in names])
    return splits, multilabel


def load_real_dataset(name: str) -> "BaseGraph":
    """datasets.py:127-227 for ppi_bp / hpo_metab / hpo_neuro / em_user (files are not shipped with the
    reference: README.md:26 gives a download link).  Same conventions: validation and test splits swap when the
    validation split is the smaller one (:168-171), mask 0/1/2 in train/val/test order, single-label targets as
    a float vector, multi-label targets as a float indicator matrix, node count = 1 + largest id seen."""
    root = real_dataset_dir(name)
    sub_path, edge_path = os.path.join(root, "subgraphs.pth"), os.path.join(root, "edge_list.txt")
    if not (os.path.exists(sub_path) and os.path.exists(edge_path)):
        raise FileNotFoundError(f"{name}: expected {sub_path} and {edge_path} (download link in the reference's "
                                f"README; set GLASS_DATASET_DIR if they live elsewhere)")
    splits, multilabel = parse_subgraph_file(sub_path)
    order = ["train", "val", "test"]
    if len(splits["val"][0]) < len(splits["test"][0]):
        order = ["train", "test", "val"]
    rows = [r for k in order for r in splits[k][0]]
    labels = [l for k in order for l in splits[k][1]]
    mask = torch.cat([torch.full((len(splits[k][0]),), i, dtype=torch.int64) for i, k in enumerate(order)])
    if multilabel:
        width = max(max(l) for l in labels) + 1
        y = torch.zeros(len(labels), width)
        for i, l in enumerate(labels):
            y[i, torch.as_tensor(l, dtype=torch.int64)] = 1
    else:
        y = torch.tensor([l[0] for l in labels], dtype=torch.float)
    pos = _pad_rows(rows)
    raw = np.loadtxt(edge_path, dtype=np.int64, ndmin=2)[:, :2]
    n = int(max(int(pos.max()), int(raw.max()))) + 1
    # the reference reads the list into an undirected networkx Graph (:220), which keeps one copy of every
    # unordered pair however often / in whichever direction the file lists it
    und = np.unique(np.minimum(raw[:, 0], raw[:, 1]) * n + np.maximum(raw[:, 0], raw[:, 1]))
    edge = torch.from_numpy(np.stack((und // n, und % n)))
    return BaseGraph(torch.empty((n, 1, 0)), edge, torch.ones(edge.shape[1]), pos, y, mask)


def pretrained_embedding(name: str, dim: int, n: int, seed: int = 0) -> torch.Tensor:
    """./Emb/<dataset>_<dim>.pt (GLASSTest.py:153-157) when the file exists, else the seeded stand-in."""
    path = os.path.join(os.environ.get("GLASS_EMB_DIR", "./Emb"), f"{name}_{dim}.pt")
    if os.path.exists(path):
        emb = torch.load(path, map_location="cpu").detach().to(torch.float32)
        if emb.shape != (n, dim):
            raise RuntimeError(f"{path}: expected shape {(n, dim)}, found {tuple(emb.shape)}")
        return emb
    return synthetic_embedding(n, dim, seed)
