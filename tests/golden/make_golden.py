"""Generate golden vectors by running the UNMODIFIED reference in the build container.

    python tests/golden/make_golden.py          # needs /root/reference (absent on the GPU box)

The reference (pure Python, /root/reference) is imported as-is with oracle/pyg_shim on
sys.path (PyG 1.7.2 is not installable offline).  Outputs are small .npz/.json fixtures
committed next to this script; tests replay them against the oracle (CPU) and against the
CUDA path (GPU).  Nothing here is imported by the product.
"""
import functools
import hashlib
import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("GLASS_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "oracle", "pyg_shim"))
sys.path.insert(0, REF)

from impl import models, utils  # noqa: E402  (the reference itself)
from impl import SubGDataset, train as ref_train  # noqa: E402


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def random_graph(n, n_und, seed, isolated=()):
    g = np.random.default_rng(seed)
    a = g.integers(0, n, size=3 * n_und)
    b = g.integers(0, n, size=3 * n_und)
    keep = a != b
    for i in isolated:
        keep &= (a != i) & (b != i)
    a, b = a[keep], b[keep]
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    key = np.unique(lo * n + hi)[:n_und]
    lo, hi = key // n, key % n
    row = np.concatenate([lo, hi])
    col = np.concatenate([hi, lo])
    order = np.lexsort((col, row))
    return np.stack([row[order], col[order]]).astype(np.int64)


def random_subgraphs(n, b, lmax, seed, lmin=2):
    g = np.random.default_rng(seed)
    pad = -np.ones((b, lmax), dtype=np.int64)
    for i in range(b):
        k = int(g.integers(lmin, lmax + 1))
        pad[i, :k] = g.choice(n, size=k, replace=False)
    pad[0, :] = g.choice(n, size=lmax, replace=False)  # at least one full row
    return pad


# ------------------------------------------------------------------------------------------
def gen_utils():
    out = {}
    pad = torch.tensor([[0, 2, 3], [1, 4, 5], [6, 7, -1]])            # impl/utils.py:9, 21
    b, p = utils.pad2batch(pad)
    out["doc_pad"], out["doc_batch"], out["doc_pos"] = pad.numpy(), b.numpy(), p.numpy()
    out["doc_z"] = utils.MaxZOZ(torch.zeros(9, 1), pad).numpy()        # impl/utils.py:33-45
    out["doc_batch2pad"] = utils.batch2pad(torch.tensor([0, 1, 0, 0, 1, 1, 2, 2])).numpy()
    rp = torch.from_numpy(random_subgraphs(500, 7, 13, 3))
    rp[3, :] = -1                                                      # an empty (all-pad) row
    b, p = utils.pad2batch(rp)
    out["rand_pad"], out["rand_batch"], out["rand_pos"] = rp.numpy(), b.numpy(), p.numpy()
    out["rand_z"] = utils.MaxZOZ(torch.zeros(500, 1), rp).numpy()
    np.savez_compressed(os.path.join(HERE, "utils_kat.npz"), **out)


def coalesced(ei, ew, n, aggr):
    adj = models.buildAdj(torch.from_numpy(ei), torch.from_numpy(ew), n, aggr).coalesce()
    return adj.indices().numpy(), adj.values().numpy()


def gen_buildadj():
    cases = {}
    g = np.random.default_rng(11)
    # (a) sorted, unique, symmetric, unit weights, one isolated node
    ei = random_graph(97, 300, 1, isolated=(5,))
    cases["sym_unit"] = (ei, np.ones(ei.shape[1], np.float32), 97)
    # (b) unsorted directed with duplicates, self loops, non-unit weights, isolated rows
    m = 400
    ei = np.stack([g.integers(0, 60, m), g.integers(0, 60, m)]).astype(np.int64)
    ei[:, :20] = ei[:, 20:40]                                          # forced duplicates
    ei[1, 50:60] = ei[0, 50:60]                                        # self loops
    ew = (g.random(m).astype(np.float32) * 3 + 0.25).astype(np.float32)
    cases["unsorted_dup_selfloop"] = (ei, ew, 64)                      # rows 60..63 isolated
    # (c) tiny weights: rows whose degree < 0.5 receive +1 (models.py:94)
    ei = random_graph(40, 80, 2)
    ew = (g.random(ei.shape[1]).astype(np.float32) * 0.05).astype(np.float32)
    cases["tiny_weights"] = (ei, ew, 40)
    # (d) single edge / empty-ish
    cases["single"] = (np.array([[2], [0]], np.int64), np.array([2.5], np.float32), 4)
    out = {}
    for name, (ei, ew, n) in cases.items():
        out[f"{name}.ei"], out[f"{name}.ew"], out[f"{name}.n"] = ei, ew, np.int64(n)
        for aggr in ("mean", "sum", "gcn"):
            idx, val = coalesced(ei, ew, n, aggr)
            out[f"{name}.{aggr}.idx"], out[f"{name}.{aggr}.val"] = idx, val
    np.savez_compressed(os.path.join(HERE, "buildadj.npz"), **out)

    # shipped graphs: digests of the reference's coalesced result (CSR form) for 3 aggr modes
    digests = {}
    for name in ("density", "cut_ratio", "coreness", "component"):
        d = np.load(os.path.join(ROOT, "data", f"{name}.npz"))
        from torch_geometric.utils import to_undirected
        e = torch.from_numpy(d["edge"].astype(np.int64))
        ei, ew = to_undirected(e, torch.ones(e.shape[1]))              # datasets.py:68-71
        n = int(d["n_node"])
        digests[name] = {"n": n, "nnz": int(ei.shape[1]), "ei_sha": sha(ei.numpy())}
        for aggr in ("mean", "sum", "gcn"):
            idx, val = coalesced(ei.numpy(), ew.numpy(), n, aggr)
            rowptr = np.zeros(n + 1, np.int64)
            np.add.at(rowptr, idx[0] + 1, 1)
            rowptr = np.cumsum(rowptr).astype(np.int32)
            digests[name][aggr] = {"rowptr": sha(rowptr), "col": sha(idx[1].astype(np.int32)),
                                   "val": sha(val.astype(np.float32)),
                                   "val_head": [float(v) for v in val[:4]]}
    with open(os.path.join(HERE, "buildadj_shipped.json"), "w") as f:
        json.dump(digests, f, indent=1, sort_keys=True)


# ------------------------------------------------------------------------------------------
POOLS = {"mean": models.MeanPool, "max": models.MaxPool, "sum": models.AddPool, "size": models.SizePool}
ACTS = {"elu": lambda: nn.ELU(inplace=True), "relu": lambda: nn.ReLU(inplace=True)}

MODEL_CASES = {
    # name: N, und.edges, H, L, aggr, pool, z_ratio, act, jk, out_dim, emb ("one"|"nodeid"), B, Lmax, use_z
    "density_like":   dict(n=257, e=900, H=8, L=1, aggr="sum", pool="size", z=1.0, act="elu", jk=1, out=3, emb="one", B=2, lmax=20, use_z=True),
    "cutratio_like":  dict(n=257, e=2000, H=8, L=1, aggr="sum", pool="mean", z=0.75, act="elu", jk=1, out=3, emb="one", B=3, lmax=20, use_z=True),
    "component_like": dict(n=301, e=1500, H=17, L=1, aggr="sum", pool="sum", z=0.9, act="elu", jk=1, out=1, emb="one", B=8, lmax=37, use_z=True),
    "coreness_like":  dict(n=257, e=3000, H=20, L=2, aggr="sum", pool="mean", z=1.0, act="elu", jk=1, out=3, emb="one", B=2, lmax=20, use_z=True),
    "ppibp_like":     dict(n=311, e=2800, H=64, L=2, aggr="mean", pool="sum", z=0.95, act="elu", jk=1, out=6, emb="nodeid", B=16, lmax=12, use_z=True),
    "emuser_like":    dict(n=311, e=6000, H=64, L=1, aggr="gcn", pool="size", z=0.75, act="elu", jk=1, out=1, emb="nodeid", B=6, lmax=60, use_z=True),
    "maxpool_relu":   dict(n=200, e=700, H=16, L=3, aggr="gcn", pool="max", z=0.8, act="relu", jk=0, out=4, emb="nodeid", B=5, lmax=9, use_z=False),
}


def build_ref_model(c, max_deg):
    """Same construction as GLASSTest.buildModel (GLASSTest.py:129-175)."""
    conv = models.EmbZGConv(c["H"], c["H"], c["L"], max_deg=max_deg, activation=ACTS[c["act"]](),
                            jk=c["jk"], dropout=0.0,
                            conv=functools.partial(models.GLASSConv, aggr=c["aggr"], z_ratio=c["z"],
                                                   dropout=0.0), gn=True)
    if c["emb"] == "nodeid":
        conv.input_emb = nn.Embedding.from_pretrained(torch.randn(c["n"], c["H"]) * 2.0, freeze=False)
    mlp = nn.Linear(c["H"] * c["L"] if c["jk"] else c["H"], c["out"])
    return models.GLASS(conv, nn.ModuleList([mlp]), nn.ModuleList([POOLS[c["pool"]]()]))


def gen_models():
    for name, c in MODEL_CASES.items():
        torch.manual_seed(1234)
        n = c["n"]
        ei = torch.from_numpy(random_graph(n, c["e"], 7, isolated=(3,)))
        ew = torch.ones(ei.shape[1])
        if c["emb"] == "one":
            x = torch.ones((n, 1, 1), dtype=torch.int64)               # datasets.py:54-56
        else:
            x = torch.arange(n, dtype=torch.int64).reshape(n, 1, 1)    # datasets.py:58-61
        pos = torch.from_numpy(random_subgraphs(n, c["B"], c["lmax"], 5))
        z = utils.MaxZOZ(x, pos) if c["use_z"] else None
        model = build_ref_model(c, int(x.max()))
        # perturb the GraphNorm parameters away from (1, 0, 1) so that they matter
        with torch.no_grad():
            for k, p in model.named_parameters():
                if "gn" in k:
                    p.add_(0.3 * torch.randn_like(p))
        sd0 = {k: v.detach().clone().numpy() for k, v in model.state_dict().items()}
        model.eval()
        with torch.no_grad():
            emb = model.NodeEmb(x, ei, ew, z)
            pooled = model.Pool(emb, pos, model.pools[0])
            logits = model(x, ei, ew, pos, z)
        # gradients in train() mode (dropout p = 0 => deterministic), real loss of GLASSTest.py:55-71
        model.train()
        if c["out"] == 1:
            y = (torch.rand(c["B"]) > 0.5).float()
            loss = nn.BCEWithLogitsLoss()(model(x, ei, ew, pos, z).flatten(), y.flatten())
        else:
            y = torch.randint(0, c["out"], (c["B"],))
            loss = nn.CrossEntropyLoss()(model(x, ei, ew, pos, z), y)
        model.zero_grad()
        loss.backward()
        out = {f"sd.{k}": v for k, v in sd0.items()}
        out.update({f"grad.{k}": p.grad.detach().numpy() for k, p in model.named_parameters()})
        out.update(ei=ei.numpy(), ew=ew.numpy(), x=x.numpy(), pos=pos.numpy(), y=y.numpy(),
                   z=(z.numpy() if z is not None else np.zeros(0, np.int64)),
                   emb=emb.numpy(), pooled=pooled.numpy(), logits=logits.numpy(),
                   loss=np.float32(loss.item()), cfg=json.dumps(c))
        np.savez_compressed(os.path.join(HERE, f"model_{name}.npz"), **out)
        print(name, "loss", float(loss), "logits", logits.flatten()[:3].tolist())


def gen_edgegnn():
    """EmbGConv(MyGCNConv) + EdgeGNN (impl/models.py:361-509; built as in GNNEmb.py:76-100)."""
    torch.manual_seed(77)
    n, H, L = 300, 32, 2
    ei = torch.from_numpy(random_graph(n, 1500, 9))
    ew = torch.ones(ei.shape[1])
    x = torch.randint(0, 12, (n, 1, 1))
    conv = models.EmbGConv(H, H, H, L, max_deg=11, activation=nn.ReLU(inplace=True), jk=True, dropout=0.0,
                           conv=functools.partial(models.MyGCNConv, aggr="mean", activation=nn.ReLU(inplace=True)), gn=True)
    mlp = nn.Linear(H * L, 1)
    model = models.EdgeGNN(conv, nn.ModuleList([mlp]), nn.ModuleList([models.MeanPool()]))
    with torch.no_grad():
        for k, p in model.named_parameters():
            if "gn" in k:
                p.add_(0.3 * torch.randn_like(p))
    sd0 = {k: v.detach().clone().numpy() for k, v in model.state_dict().items()}
    pairs = torch.randint(0, n, (40, 2))
    y = (torch.rand(40) > 0.5).float()
    model.train()
    logits = model(x, ei, ew, pairs)
    loss = nn.BCEWithLogitsLoss()(logits.flatten(), y)
    model.zero_grad()
    loss.backward()
    out = {f"sd.{k}": v for k, v in sd0.items()}
    out.update({f"grad.{k}": p.grad.detach().numpy() for k, p in model.named_parameters()})
    out.update(ei=ei.numpy(), ew=ew.numpy(), x=x.numpy(), pairs=pairs.numpy(), y=y.numpy(),
               logits=logits.detach().numpy(), loss=np.float32(loss.item()))
    np.savez_compressed(os.path.join(HERE, "model_edgegnn.npz"), **out)
    print("edgegnn loss", float(loss))


# ------------------------------------------------------------------------------------------
def gen_density_trajectory(n_steps=62, nodeid=False):
    """Seed-0 loss trajectory of the unmodified reference on the shipped density config
    (config/density.yml; GLASSTest.py flags --use_one --use_seed --use_maxzeroone --repeat 1).
    Reproduces GLASSTest.py:34-47, 77-126, 129-175, 205-216 without its argparse side effects."""
    import random
    import yaml
    import datasets as ref_datasets
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        with open("config/density.yml") as f:
            params = yaml.safe_load(f)

        def set_seed(seed):
            random.seed(seed)
            np.random.seed(seed)
            torch.manual_seed(seed)

        set_seed(0)
        ref_datasets.load_dataset("density")           # GLASSTest.py:49 (consumes one randperm)
        ref_datasets.load_dataset("density")           # split() at GLASSTest.py:278
        set_seed(0)                                    # (1 << 0) - 1, GLASSTest.py:205
        baseG = ref_datasets.load_dataset("density")   # split() at GLASSTest.py:207
        baseG.y = baseG.y.to(torch.int64)
        if nodeid:
            baseG.setNodeIdFeature()
        else:
            baseG.setOneFeature()
        max_deg = torch.max(baseG.x)
        trn = SubGDataset.GDataset(*baseG.get_split("train"))
        c = dict(H=params["hidden_dim"], L=params["conv_layer"], aggr=params["aggr"], z=params["z_ratio"],
                 act="elu", jk=1, out=3, emb="one", pool=params["pool"], n=baseG.x.shape[0])
        conv = models.EmbZGConv(c["H"], c["H"], c["L"], max_deg=max_deg, activation=nn.ELU(inplace=True),
                                jk=1, dropout=params["dropout"],
                                conv=functools.partial(models.GLASSConv, aggr=c["aggr"], z_ratio=c["z"],
                                                       dropout=params["dropout"]), gn=True)
        if nodeid:  # GLASSTest.py:153-157 with a seeded synthetic table (no Emb/density_8.pt is shipped)
            table = torch.randn(baseG.x.shape[0], c["H"], generator=torch.Generator().manual_seed(99)) * 2.0
            conv.input_emb = nn.Embedding.from_pretrained(table, freeze=False)
        mlp = nn.Linear(c["H"] * c["L"], 3)
        gnn = models.GLASS(conv, nn.ModuleList([mlp]), nn.ModuleList([POOLS[params["pool"]]()]))
        sd0 = {k: v.detach().clone().numpy() for k, v in gnn.state_dict().items()}
        loader = SubGDataset.ZGDataloader(trn, params["batch_size"], z_fn=utils.MaxZOZ, shuffle=True,
                                          drop_last=True)
        opt = torch.optim.Adam(gnn.parameters(), lr=params["lr"])
        loss_fn = nn.CrossEntropyLoss()
        gnn.train()
        losses, batches, ys = [], [], []
        for batch in loader:
            opt.zero_grad()
            pred = gnn(*batch[:-1], id=0)
            loss = loss_fn(pred, batch[-1])
            loss.backward()
            losses.append(loss.item())
            opt.step()
            batches.append(batch[3].numpy().copy())
            ys.append(batch[-1].numpy().copy())
            if len(losses) >= n_steps:
                break
        out = {f"sd.{k}": v for k, v in sd0.items()}
        if not nodeid:
            out.update({f"sd_end.{k}": v.detach().numpy() for k, v in gnn.state_dict().items()})
        out.update(losses=np.array(losses, np.float64), pos=np.stack(batches), y=np.stack(ys),
                   mask=baseG.mask.numpy(), params=json.dumps(params))
        np.savez_compressed(os.path.join(HERE, "trajectory_density_nodeid.npz" if nodeid else
                                         "trajectory_density.npz"), **out)
        print("density trajectory", losses[:3], "...", losses[-1], "steps", len(losses))
    finally:
        os.chdir(cwd)


if __name__ == "__main__":
    gen_utils()
    gen_buildadj()
    gen_models()
    gen_edgegnn()
    gen_density_trajectory()
    gen_density_trajectory(n_steps=40, nodeid=True)
    print("golden fixtures:", sorted(f for f in os.listdir(HERE) if f.endswith((".npz", ".json"))))
