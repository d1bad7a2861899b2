"""impl/SubGDataset.py equivalents: split container and the batch loaders that feed train/test.

A batch is the tuple (x, edge_index, edge_attr, subG_node[perm], z, y[perm]) (impl/SubGDataset.py:92-96):
the whole base graph plus the padded node sets, labels z and targets of the sampled subgraphs.
Index sampling goes through torch.utils.data.DataLoader exactly like the reference, so a seeded run
draws the same permutations.
"""
from __future__ import annotations

import torch
from torch.utils.data import DataLoader


class GDataset:
    def __init__(self, x, edge_index, edge_attr, pos, y):
        self.x, self.edge_index, self.edge_attr, self.pos, self.y = x, edge_index, edge_attr, pos, y
        self.num_nodes = x.shape[0]

    def __len__(self):
        return self.pos.shape[0]

    def __getitem__(self, idx):
        return self.pos[idx], self.y[idx]

    def to(self, device):
        for name in ("x", "edge_index", "edge_attr", "pos", "y"):
            setattr(self, name, getattr(self, name).to(device))
        return self


class GDataloader(DataLoader):
    def __init__(self, Gdataset, batch_size=64, shuffle=True, drop_last=False):
        super().__init__(torch.arange(len(Gdataset)).to(Gdataset.x.device), batch_size=batch_size, shuffle=shuffle,
                         drop_last=drop_last)
        self.Gdataset = Gdataset

    def get_x(self):
        return self.Gdataset.x

    def get_ei(self):
        return self.Gdataset.edge_index

    def get_ea(self):
        return self.Gdataset.edge_attr

    def get_pos(self):
        return self.Gdataset.pos

    def get_y(self):
        return self.Gdataset.y

    def __iter__(self):
        self.iter = super().__iter__()
        return self

    def __next__(self):
        perm = next(self.iter)
        return self.get_x(), self.get_ei(), self.get_ea(), self.get_pos()[perm], self.get_y()[perm]


class ZGDataloader(GDataloader):
    """Adds the per-batch node labels z = z_fn(x, subG_node[perm]) (MaxZOZ for --use_maxzeroone)."""

    def __init__(self, Gdataset, batch_size=64, shuffle=True, drop_last=False,
                 z_fn=lambda x, y: torch.zeros((x.shape[0], x.shape[1]), dtype=torch.int64)):
        super().__init__(Gdataset, batch_size, shuffle, drop_last)
        self.z_fn = z_fn

    def __next__(self):
        perm = next(self.iter)
        tpos = self.get_pos()[perm]
        return self.get_x(), self.get_ei(), self.get_ea(), tpos, self.z_fn(self.get_x(), tpos), self.get_y()[perm]


def index_batches(loader: GDataloader):
    """The index batches `iter(loader)` would produce, as CPU int64 tensors, drawing from torch's global RNG
    in the same order as torch.utils.data.DataLoader does (the iterator's base seed first, then the
    RandomSampler's own seed on its first draw), so a seeded run sees the same subgraph order whichever
    way the epoch is iterated.  tests/test_host_logic.py pins this against DataLoader itself."""
    if loader.generator is None:
        torch.empty((), dtype=torch.int64).random_()          # _BaseDataLoaderIter.__init__: self._base_seed
    else:
        torch.empty((), dtype=torch.int64).random_(generator=loader.generator)
    for idx in loader.batch_sampler:
        yield torch.as_tensor(idx, dtype=torch.int64)


def epoch_batches(loader: GDataloader):
    """(subG_node, y) of every batch of one epoch in the loader's order, for the captured train / eval
    steps: one gather of the whole epoch instead of one per batch, no per-batch z_fn (the captured step
    computes the labels itself), no per-batch collation of device scalars."""
    batches = list(index_batches(loader))
    if not batches:
        return
    pos, y = loader.get_pos(), loader.get_y()
    order = torch.cat(batches).to(pos.device, non_blocking=True)
    pos_e, y_e = pos[order], y[order]
    off = 0
    for b in batches:
        n = b.numel()
        yield pos_e[off:off + n], y_e[off:off + n]
        off += n
