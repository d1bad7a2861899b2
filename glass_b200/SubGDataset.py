"""Split container and batch loaders with the public names of impl/SubGDataset.py.

A batch is `(x, edge_index, edge_attr, subG_node[idx], [z,] y[idx])` (impl/SubGDataset.py:92-96): the whole
base graph by reference plus the padded node sets and targets of the sampled subgraphs.  The loaders here
are plain iterables built on torch's samplers; they draw from torch's global RNG in exactly the order
`torch.utils.data.DataLoader` does (the reference subclasses it), so a seeded run samples the same
subgraphs in the same order -- `tests/test_host_logic.py` pins that against DataLoader itself.
"""
from __future__ import annotations

from typing import Callable, Iterator, Tuple

import torch
from torch.utils.data import BatchSampler, RandomSampler, SequentialSampler

_FIELDS = ("x", "edge_index", "edge_attr", "pos", "y")


class GDataset:
    """One split: base graph (x, edge_index, edge_attr) + padded subgraph node sets `pos` [S, Lmax] (-1 = pad)
    + targets `y` [S]  (impl/SubGDataset.py:6-36)."""

    def __init__(self, x, edge_index, edge_attr, pos, y):
        for name, value in zip(_FIELDS, (x, edge_index, edge_attr, pos, y)):
            setattr(self, name, value)
        self.num_nodes = x.shape[0]
        # the reference fails with an index error when a subgraph names a node outside the graph (emb[pos],
        # impl/models.py:348); the pooling / label kernels skip such ids, so the split is checked once here
        if pos.numel() and (int(pos.max()) >= self.num_nodes or int(pos.min()) < -1):
            raise IndexError(f"subG_node holds node ids outside [-1, {self.num_nodes})")

    def __len__(self) -> int:
        return self.pos.shape[0]

    def __getitem__(self, idx):
        return self.pos[idx], self.y[idx]

    def to(self, device):
        for name in _FIELDS:
            setattr(self, name, getattr(self, name).to(device))
        return self


def _zero_labels(x, pos):
    return torch.zeros((x.shape[0], x.shape[1]), dtype=torch.int64)


class GDataloader:
    """Iterable over the subgraph batches of a GDataset (impl/SubGDataset.py:38-73).

    `generator=None` means torch's global RNG, like DataLoader's default."""

    def __init__(self, Gdataset: GDataset, batch_size: int = 64, shuffle: bool = True, drop_last: bool = False,
                 generator=None):
        self.Gdataset, self.batch_size, self.generator = Gdataset, batch_size, generator
        ids = range(len(Gdataset))
        sampler = RandomSampler(ids, generator=generator) if shuffle else SequentialSampler(ids)
        self.batch_sampler = BatchSampler(sampler, batch_size, drop_last)

    def __len__(self) -> int:
        return len(self.batch_sampler)

    def index_batches(self) -> Iterator[torch.Tensor]:
        """CPU int64 index batches of one epoch.  RNG order of DataLoader: the iterator's base seed is drawn
        when iteration starts, the RandomSampler's own seed when the first index is requested."""
        torch.empty((), dtype=torch.int64).random_(generator=self.generator)
        for idx in self.batch_sampler:
            yield torch.as_tensor(idx, dtype=torch.int64)

    def _batch(self, pos, y) -> Tuple:
        ds = self.Gdataset
        return ds.x, ds.edge_index, ds.edge_attr, pos, y

    def __iter__(self):
        ds = self.Gdataset
        for idx in self.index_batches():
            idx = idx.to(ds.pos.device)
            yield self._batch(ds.pos[idx], ds.y[idx])


# accessor methods of the reference loader (impl/SubGDataset.py:50-63)
for _method, _field in zip(("get_x", "get_ei", "get_ea", "get_pos", "get_y"), _FIELDS):
    setattr(GDataloader, _method, (lambda field: lambda self: getattr(self.Gdataset, field))(_field))


class ZGDataloader(GDataloader):
    """Adds the per-batch node labels z = z_fn(x, subG_node[idx]) before the target (impl/SubGDataset.py:76-96);
    `utils.MaxZOZ` for --use_maxzeroone, all-zero labels by default."""

    def __init__(self, Gdataset: GDataset, batch_size: int = 64, shuffle: bool = True, drop_last: bool = False,
                 z_fn: Callable = _zero_labels, generator=None):
        super().__init__(Gdataset, batch_size, shuffle, drop_last, generator)
        self.z_fn = z_fn

    def _batch(self, pos, y) -> Tuple:
        ds = self.Gdataset
        return ds.x, ds.edge_index, ds.edge_attr, pos, self.z_fn(ds.x, pos), y


def index_batches(loader: GDataloader) -> Iterator[torch.Tensor]:
    return loader.index_batches()


def epoch_batches(loader: GDataloader):
    """(subG_node, y) of every batch of one epoch in the loader's order, for the captured train / eval
    steps: one gather of the whole epoch instead of one per batch, no per-batch z_fn (the captured step
    computes the labels itself)."""
    batches = list(loader.index_batches())
    if not batches:
        return
    ds = loader.Gdataset
    order = torch.cat(batches).to(ds.pos.device, non_blocking=True)
    pos_e, y_e = ds.pos[order], ds.y[order]
    off = 0
    for b in batches:
        n = b.numel()
        yield pos_e[off:off + n], y_e[off:off + n]
        off += n
