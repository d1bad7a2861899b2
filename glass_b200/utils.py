"""impl/utils.py equivalents: batch2pad (host helper), pad2batch and MaxZOZ (GPU kernels)."""
from __future__ import annotations

import torch

from . import ops


def batch2pad(batch: torch.Tensor) -> torch.Tensor:
    """impl/utils.py:5-15: batch [0,1,0,0,1,1,2,2] -> pad [[0,2,3],[1,4,5],[6,7,-1]].  Host-side data
    preparation only (not on the training path)."""
    ids = torch.unique(batch)
    ids = ids[ids >= 0]
    pos = torch.arange(batch.shape[0], device=batch.device)
    rows = [pos[batch == i] for i in ids]
    width = max((r.numel() for r in rows), default=0)
    pad = torch.full((len(rows), width), -1, dtype=torch.int64, device=batch.device)
    for i, r in enumerate(rows):
        pad[i, :r.numel()] = r
    return pad


def pad2batch(pad: torch.Tensor):
    """impl/utils.py:18-29: (batch_ids, node_ids) of the entries >= 0 of a padded matrix, row-major."""
    return ops.pad2batch(pad)


def MaxZOZ(x: torch.Tensor, pos: torch.Tensor) -> torch.Tensor:
    """impl/utils.py:32-45: max-zero-one labels, int64 [N]; z[n] = 1 iff node n is in any row of pos."""
    return ops.maxzoz(x.shape[0], pos.to(x.device))
