"""Model / task construction mirroring GLASSTest.py (buildModel :129-175, loss & score :55-71,
hyper-parameters of the reference's config/<dataset>.yml), for scripts, benchmarks and tests."""
from __future__ import annotations

import functools
import os
from typing import Optional

import torch
import torch.nn as nn
import yaml

from . import metrics, models

_CONFIG_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "config")
# which dataset's row the synthetic shapes borrow their hyper-parameters from
CONFIG_OF = {"ppi_bp_shaped": "ppi_bp", "em_user_shaped": "em_user", "em_user_shaped_powerlaw": "em_user",
             "stress": "em_user", "stress_small": "em_user"}


def load_params(dataset: str) -> dict:
    """Hyper-parameters of a dataset: config/<dataset>.yml, the reference's own files (GLASSTest.py:273-276)."""
    name = CONFIG_OF.get(dataset, dataset)
    own = os.path.join(_CONFIG_DIR, f"{name}.yml")
    if not os.path.exists(own):
        raise FileNotFoundError(f"no hyper-parameters for dataset {dataset!r} ({own})")
    with open(own) as f:
        return yaml.safe_load(f)


def build_model(hidden_dim, conv_layer, dropout, jk, pool, z_ratio, aggr, max_deg, output_channels,
                pretrained: Optional[torch.Tensor] = None, device=None):
    """GLASSTest.buildModel (GLASSTest.py:129-175): EmbZGConv of GLASSConv layers + Linear head + pool."""
    conv = models.EmbZGConv(hidden_dim, hidden_dim, conv_layer, max_deg=max_deg, activation=nn.ELU(inplace=True),
                            jk=jk, dropout=dropout,
                            conv=functools.partial(models.GLASSConv, aggr=aggr, z_ratio=z_ratio, dropout=dropout),
                            gn=True)
    if pretrained is not None:  # --use_nodeid, GLASSTest.py:153-157
        conv.input_emb = nn.Embedding.from_pretrained(pretrained, freeze=False)
    mlp = nn.Linear(hidden_dim * conv_layer if jk else hidden_dim, output_channels)
    pool_cls = {"mean": models.MeanPool, "max": models.MaxPool, "sum": models.AddPool, "size": models.SizePool}
    if pool not in pool_cls:
        raise NotImplementedError(pool)  # GLASSTest.py:171
    gnn = models.GLASS(conv, nn.ModuleList([mlp]), nn.ModuleList([pool_cls[pool]()]))
    return gnn.to(device) if device is not None else gnn


def task_of(y: torch.Tensor):
    """GLASSTest.py:55-71: (loss_fn, output_channels, score_fn, y in the dtype the loss needs)."""
    if y.unique().shape[0] == 2:
        def loss_fn(x, t):
            return nn.BCEWithLogitsLoss()(x.flatten(), t.flatten())
        y = y.to(torch.float)
        return loss_fn, (y.shape[1] if y.ndim > 1 else 1), metrics.binaryf1, y
    y = y.to(torch.int64)
    return nn.CrossEntropyLoss(), int(y.unique().shape[0]), metrics.microf1, y
