"""Epoch drivers with the call signatures of impl/train.py (train :4-17, test :20-34).

A loader batch is a tuple whose last element is the target and whose other elements are the positional
arguments of the model (impl/SubGDataset.py:92-96).  The captured-graph equivalents live in graphed.py
(train_epoch / test_epoch); these eager versions are what `GLASSTest.py` uses without --graph.
"""
from __future__ import annotations

import torch


def train(optimizer, model, dataloader, loss_fn, sync_each_step: bool = True):
    """One optimizer step per batch; returns the mean of the per-batch losses (impl/train.py:17).

    sync_each_step=True reads every loss back right away, which is what the reference's `.item()`
    (impl/train.py:15) does: one host sync per step.  False keeps the losses on the device until the
    epoch ends -- same value, no per-step sync."""
    model.train()
    history = []
    for *inputs, target in dataloader:
        optimizer.zero_grad()
        step_loss = loss_fn(model(*inputs, id=0), target)
        step_loss.backward()
        optimizer.step()
        history.append(step_loss.detach())
        if sync_each_step:
            history[-1] = history[-1].item()
    if sync_each_step:
        return sum(history) / len(history)
    return float(torch.stack(history).double().sum().item()) / len(history)


@torch.no_grad()
def test(model, dataloader, metrics, loss_fn):
    """(metrics(logits, targets), loss_fn(logits, targets)) over the whole loader, eval mode (impl/train.py:34)."""
    model.eval()
    pairs = [(model(*inputs), target) for *inputs, target in dataloader]
    logits = torch.cat([out for out, _ in pairs], dim=0)
    targets = torch.cat([t for _, t in pairs], dim=0)
    return metrics(logits.cpu().numpy(), targets.cpu().numpy()), loss_fn(logits, targets)
