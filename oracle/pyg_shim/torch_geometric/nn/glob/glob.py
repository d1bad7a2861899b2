import torch


def _size(batch, size):
    return int(batch.max()) + 1 if size is None else size


def global_add_pool(x, batch, size=None):
    out = torch.zeros(_size(batch, size), x.shape[1], dtype=x.dtype, device=x.device)
    return out.index_add_(0, batch, x)


def global_mean_pool(x, batch, size=None):
    n = _size(batch, size)
    cnt = torch.zeros(n, dtype=x.dtype, device=x.device)
    cnt.index_add_(0, batch, torch.ones_like(batch, dtype=x.dtype))
    return global_add_pool(x, batch, n) / cnt.clamp(min=1).unsqueeze(-1)


def global_max_pool(x, batch, size=None):
    n = _size(batch, size)
    out = torch.zeros(n, x.shape[1], dtype=x.dtype, device=x.device)
    idx = batch.view(-1, 1).expand_as(x)
    return out.scatter_reduce(0, idx, x, reduce="amax", include_self=False)
