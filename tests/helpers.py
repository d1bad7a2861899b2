"""Shared helpers for the parity tests (test infrastructure)."""
import json
import os

import numpy as np
import torch

from oracle import glass_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODEL_CASES = ["density_like", "cutratio_like", "component_like", "coreness_like", "ppibp_like",
               "emuser_like", "maxpool_relu"]


def load_model_case(name):
    d = np.load(os.path.join(GOLDEN, f"model_{name}.npz"))
    c = json.loads(str(d["cfg"]))
    cfg = O.GlassConfig(hidden_dim=c["H"], conv_layer=c["L"], aggr=c["aggr"], z_ratio=c["z"], dropout=0.0,
                        pool=c["pool"], jk=bool(c["jk"]), activation=c["act"], out_dim=c["out"])
    sd = {k[3:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("sd.")}
    grads = {k[5:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("grad.")}
    t = lambda k: torch.from_numpy(d[k])
    z = t("z") if d["z"].shape[0] else None
    return dict(cfg=cfg, raw=c, sd=sd, grads=grads, ei=t("ei"), ew=t("ew"), x=t("x"), pos=t("pos"), y=t("y"),
                z=z, emb=t("emb"), pooled=t("pooled"), logits=t("logits"), loss=float(d["loss"]))


def rel_err(a, b):
    """Norm-wise relative error max|a-b| / max|b| (the 1e-4 bar of BASELINE.json's north_star)."""
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    denom = b.abs().max().clamp(min=1e-30)
    return float((a - b).abs().max() / denom)


# ------------------------------------------------------------------------------------------
# building the product model (glass_b200.models) for a golden / synthetic case
# ------------------------------------------------------------------------------------------
def build_product_model(raw, n_nodes, device=None, dropout=0.0):
    """Same construction as GLASSTest.buildModel (GLASSTest.py:129-175) on glass_b200.models."""
    import functools

    import torch.nn as nn

    from glass_b200 import models
    act = {"elu": nn.ELU(inplace=True), "relu": nn.ReLU(inplace=True)}[raw["act"]]
    max_deg = 1 if raw["emb"] == "one" else n_nodes - 1
    conv = models.EmbZGConv(raw["H"], raw["H"], raw["L"], max_deg=max_deg, activation=act, jk=raw["jk"],
                            dropout=dropout,
                            conv=functools.partial(models.GLASSConv, aggr=raw["aggr"], z_ratio=raw["z"],
                                                   dropout=dropout), gn=True)
    if raw["emb"] == "nodeid":
        conv.input_emb = nn.Embedding.from_pretrained(torch.randn(n_nodes, raw["H"]) * 2.0, freeze=False)
    mlp = nn.Linear(raw["H"] * raw["L"] if raw["jk"] else raw["H"], raw["out"])
    pool = {"mean": models.MeanPool, "max": models.MaxPool, "sum": models.AddPool, "size": models.SizePool}[raw["pool"]]()
    m = models.GLASS(conv, nn.ModuleList([mlp]), nn.ModuleList([pool]))
    return m.to(device) if device is not None else m


def keep_masks_for(raw, n, p, seed, device="cpu"):
    """Dropout keep-masks in the consumption order documented in oracle.emb_zg_conv."""
    g = torch.Generator().manual_seed(seed)
    H, L = raw["H"], raw["L"]
    count = 1 + 2 * (L - 1) + 1
    return [(torch.rand(n, H, generator=g) >= p).to(torch.uint8).to(device) for _ in range(count)]
