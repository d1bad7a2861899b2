"""Label-batch data parallelism: the base graph, its CSR and the parameters are
replicated on every GPU; rank r processes label batches r, r+P, ... of a permutation shared by all
ranks; gradients are averaged once per step (graphed.GradAverager: the table gradient in place, the small
tensors through one flat buffer).

The reference has no distributed code at all -- this is new capability, and because
max-zero-one labels are per batch (impl/utils.py:40-44) P ranks = P independent batches per step
(effective batch x P), not a re-sharding of one batch.
Backend: NCCL over NVLink on GPUs; gloo for the CPU tests of the host logic.
"""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


def shard_batches(n_batches: int, rank: int, world: int) -> List[int]:
    """Batch indices of one epoch handled by `rank`: r, r+P, ... (all ranks take the same count so
    that every step has a matching all-reduce; the remainder is dropped like drop_last)."""
    per = n_batches // world
    return [rank + i * world for i in range(per)]


def shared_permutation(n: int, seed: int, epoch: int) -> torch.Tensor:
    """The epoch's subgraph order, identical on every rank (seeded; no communication needed)."""
    g = torch.Generator().manual_seed(seed * 1000003 + epoch)
    return torch.randperm(n, generator=g)


@torch.no_grad()
def sharded_test(model, dataloader, metrics, loss_fn, group=None):
    """impl/train.py:20-34 with the loader's batches dealt round-robin to the ranks (SURVEY.md section 8e:
    "shard the eval batches, all-gather the logits").  Evaluation is per batch and deterministic (labels are
    per batch, no dropout), so every rank returns what train.test returns on one device; the logits are
    put back in loader order before the metric and the loss are taken."""
    on = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if on else 1
    rank = dist.get_rank(group) if on else 0
    batches = list(dataloader)
    model.eval()
    mine = [i for i in range(len(batches)) if i % world == rank]
    preds = [model(*batches[i][:-1]) for i in mine]
    y = torch.cat([b[-1] for b in batches], dim=0)
    if world == 1:
        pred = torch.cat(preds, dim=0)
        return metrics(pred.cpu().numpy(), y.cpu().numpy()), loss_fn(pred, y)
    rows = [int(b[-1].shape[0]) for b in batches]
    dev = y.device
    width = torch.tensor([preds[0].shape[1] if preds else 0], device=dev)
    dist.all_reduce(width, op=dist.ReduceOp.MAX, group=group)
    width = int(width)
    per_rank = [sum(rows[i] for i in range(len(batches)) if i % world == r) for r in range(world)]
    buf = torch.zeros((max(per_rank), width), dtype=torch.float32, device=dev)
    if preds:
        buf[:per_rank[rank]] = torch.cat(preds, dim=0)
    gathered = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf, group=group)
    offs = [0] * world
    parts = []
    for i, n in enumerate(rows):                       # back to loader order
        r = i % world
        parts.append(gathered[r][offs[r]:offs[r] + n])
        offs[r] += n
    pred = torch.cat(parts, dim=0)
    return metrics(pred.cpu().numpy(), y.cpu().numpy()), loss_fn(pred, y)
