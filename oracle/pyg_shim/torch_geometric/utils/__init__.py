import torch


def _coalesce(edge_index, edge_attr, n):
    key = edge_index[0] * n + edge_index[1]
    uniq, inv = torch.unique(key, sorted=True, return_inverse=True)
    ei = torch.stack((torch.div(uniq, n, rounding_mode="floor"), uniq % n))
    if edge_attr is None:
        return ei, None
    ea = torch.zeros((uniq.numel(),) + tuple(edge_attr.shape[1:]), dtype=edge_attr.dtype)
    ea.index_add_(0, inv, edge_attr)
    return ei, ea


def _num_nodes(edge_index, num_nodes):
    return int(edge_index.max()) + 1 if num_nodes is None else int(num_nodes)


def to_undirected(edge_index, edge_attr=None, num_nodes=None, reduce="add"):
    n = _num_nodes(edge_index, num_nodes)
    row, col = edge_index
    ei = torch.stack((torch.cat((row, col)), torch.cat((col, row))))
    ea = None if edge_attr is None else torch.cat((edge_attr, edge_attr))
    ei, ea = _coalesce(ei, ea, n)
    return ei if edge_attr is None else (ei, ea)


def is_undirected(edge_index, edge_attr=None, num_nodes=None):
    n = _num_nodes(edge_index, num_nodes)
    ei, _ = _coalesce(edge_index, None, n)
    return ei.size(1) == to_undirected(edge_index, num_nodes=n).size(1) and ei.size(1) == edge_index.size(1)


def negative_sampling(*a, **k):
    raise NotImplementedError("outside the GLASS hot path")


def to_networkx(*a, **k):
    raise NotImplementedError("outside the GLASS hot path")
