import torch
import torch.nn as nn


def _segment_mean(x, batch, size):
    out = torch.zeros(size, x.shape[1], dtype=x.dtype, device=x.device)
    out.index_add_(0, batch, x)
    cnt = torch.zeros(size, dtype=x.dtype, device=x.device)
    cnt.index_add_(0, batch, torch.ones_like(batch, dtype=x.dtype))
    return out / cnt.clamp(min=1).unsqueeze(-1)


class GraphNorm(nn.Module):
    """x -> weight * (x - mean_scale*mean) / sqrt(var + eps) + bias, per graph segment."""

    def __init__(self, in_channels, eps=1e-5):
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

    def forward(self, x, batch=None):
        if batch is None:
            batch = x.new_zeros(x.size(0), dtype=torch.long)
        size = int(batch.max()) + 1
        mean = _segment_mean(x, batch, size)
        out = x - mean.index_select(0, batch) * self.mean_scale
        var = _segment_mean(out * out, batch, size)
        std = (var + self.eps).sqrt().index_select(0, batch)
        return self.weight * out / std + self.bias


class GraphSizeNorm(nn.Module):
    def forward(self, x, batch=None):
        if batch is None:
            batch = x.new_zeros(x.size(0), dtype=torch.long)
        size = int(batch.max()) + 1
        cnt = torch.zeros(size, dtype=x.dtype, device=x.device)
        cnt.index_add_(0, batch, torch.ones_like(batch, dtype=x.dtype))
        inv_sqrt = cnt.pow(-0.5)
        return x * inv_sqrt.index_select(0, batch).view(-1, 1)
