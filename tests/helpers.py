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
