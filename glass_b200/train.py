"""impl/train.py equivalents: one training epoch and one evaluation pass over a loader."""
from __future__ import annotations

import torch


def train(optimizer, model, dataloader, loss_fn, sync_each_step: bool = True):
    """impl/train.py:4-17: mean loss over the epoch; one optimizer step per batch.

    sync_each_step=True reads the loss back every step exactly like the reference (`.item()`,
    impl/train.py:15 -- one host sync per step); False defers the read to the end of the epoch
    (same values, no per-step sync)."""
    model.train()
    losses = []
    for batch in dataloader:
        optimizer.zero_grad()
        pred = model(*batch[:-1], id=0)
        loss = loss_fn(pred, batch[-1])
        loss.backward()
        losses.append(loss.detach().item() if sync_each_step else loss.detach())
        optimizer.step()
    if sync_each_step:
        return sum(losses) / len(losses)
    return float(torch.stack(losses).double().sum().item()) / len(losses)


@torch.no_grad()
def test(model, dataloader, metrics, loss_fn):
    """impl/train.py:20-34: score and loss over all batches of the loader."""
    model.eval()
    preds, ys = [], []
    for batch in dataloader:
        preds.append(model(*batch[:-1]))
        ys.append(batch[-1])
    pred = torch.cat(preds, dim=0)
    y = torch.cat(ys, dim=0)
    return metrics(pred.cpu().numpy(), y.cpu().numpy()), loss_fn(pred, y)
