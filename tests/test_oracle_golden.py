"""CPU: pin the oracle against golden vectors produced by the unmodified reference."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import glass_oracle as O
from tests.helpers import GOLDEN, MODEL_CASES, load_model_case, rel_err

ROOT = os.path.dirname(GOLDEN.rstrip("/")).rsplit("/tests", 1)[0]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_utils_known_answers():
    d = np.load(os.path.join(GOLDEN, "utils_kat.npz"))
    # docstring examples of impl/utils.py:9,21,36-38
    assert d["doc_batch"].tolist() == [0, 0, 0, 1, 1, 1, 2, 2]
    assert d["doc_pos"].tolist() == [0, 2, 3, 1, 4, 5, 6, 7]
    for tag, n in (("doc", 9), ("rand", 500)):
        pad = torch.from_numpy(d[f"{tag}_pad"])
        b, p = O.pad2batch(pad)
        assert np.array_equal(b.numpy(), d[f"{tag}_batch"])
        assert np.array_equal(p.numpy(), d[f"{tag}_pos"])
        assert np.array_equal(O.max_zero_one(n, pad).numpy(), d[f"{tag}_z"])


@pytest.mark.parametrize("case", ["sym_unit", "unsorted_dup_selfloop", "tiny_weights", "single"])
@pytest.mark.parametrize("aggr", ["mean", "sum", "gcn"])
def test_build_csr_matches_reference_coalesced(case, aggr):
    d = np.load(os.path.join(GOLDEN, "buildadj.npz"))
    ei, ew, n = d[f"{case}.ei"], d[f"{case}.ew"], int(d[f"{case}.n"])
    csr = O.build_csr_numpy(ei, ew, n, aggr)
    idx, val = d[f"{case}.{aggr}.idx"], d[f"{case}.{aggr}.val"]
    rows = np.repeat(np.arange(n), np.diff(csr["rowptr"]))
    assert np.array_equal(rows, idx[0]) and np.array_equal(csr["col"], idx[1])
    if case == "unsorted_dup_selfloop" or case == "tiny_weights":
        # non-unit weights: degree / duplicate sums depend on fp32 summation order;
        # the reference's own order is an ATen implementation detail -> few-ulp bound
        np.testing.assert_allclose(csr["val"], val, rtol=1e-6, atol=0)
    else:
        assert np.array_equal(csr["val"].view(np.uint32), val.view(np.uint32))
    # the restated COO path is the same ATen op sequence as the reference -> bit-exact always
    adj = O.build_adj(torch.from_numpy(ei), torch.from_numpy(ew), n, aggr).coalesce()
    assert np.array_equal(adj.indices().numpy(), idx)
    assert np.array_equal(adj.values().numpy().view(np.uint32), val.view(np.uint32))
    # transposed CSR is the same matrix
    dense = np.zeros((n, n), np.float64)
    dense[rows, csr["col"]] = csr["val"]
    rows_t = np.repeat(np.arange(n), np.diff(csr["rowptr_t"]))
    dense_t = np.zeros((n, n), np.float64)
    dense_t[rows_t, csr["col_t"]] = csr["val_t"]
    assert np.array_equal(dense.T, dense_t)


def test_build_csr_shipped_graphs_bit_exact():
    with open(os.path.join(GOLDEN, "buildadj_shipped.json")) as f:
        dig = json.load(f)
    from glass_b200 import datasets
    for name, ref in dig.items():
        ei, ew, n = datasets.load_edges(name)
        assert sha(ei.numpy()) == ref["ei_sha"] and n == ref["n"]
        for aggr in ("mean", "sum", "gcn"):
            csr = O.build_csr_numpy(ei.numpy(), ew.numpy(), n, aggr)
            assert sha(csr["rowptr"]) == ref[aggr]["rowptr"]
            assert sha(csr["col"]) == ref[aggr]["col"]
            assert sha(csr["val"]) == ref[aggr]["val"], (name, aggr)


@pytest.mark.parametrize("name", MODEL_CASES)
def test_oracle_forward_and_grads_match_reference(name):
    c = load_model_case(name)
    cfg = c["cfg"]
    sd = {k: v.clone().requires_grad_(True) for k, v in c["sd"].items()}
    adj = O.build_adj(c["ei"], c["ew"], c["x"].shape[0], cfg.aggr)
    with torch.no_grad():
        logits, pooled, emb = O.glass_forward(sd, c["x"], adj, c["pos"], c["z"], cfg, training=False)
    assert rel_err(emb, c["emb"]) < 1e-6
    assert rel_err(pooled, c["pooled"]) < 1e-6
    assert rel_err(logits, c["logits"]) < 1e-6
    logits, _, _ = O.glass_forward(sd, c["x"], adj, c["pos"], c["z"], cfg, training=True)
    loss = O.loss_fn_for(cfg.out_dim == 1)(logits, c["y"])
    loss.backward()
    assert abs(float(loss) - c["loss"]) < 1e-6 * max(1.0, abs(c["loss"]))
    for k, g in c["grads"].items():
        assert rel_err(sd[k].grad, g) < (1e-4 if "input_emb" in k else 1e-5), k


def _replay(fname, nodeid, steps, tol):
    d = np.load(os.path.join(GOLDEN, fname))
    params = json.loads(str(d["params"]))
    from glass_b200 import datasets
    ei, ew, n = datasets.load_edges("density")
    cfg = O.GlassConfig(hidden_dim=params["hidden_dim"], conv_layer=params["conv_layer"], aggr=params["aggr"],
                        z_ratio=params["z_ratio"], dropout=params["dropout"], pool=params["pool"], jk=True,
                        activation="elu", out_dim=3)
    sd = {k[3:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("sd.")}
    m = O.OracleModel(cfg, sd)
    opt = torch.optim.Adam(m.params, lr=params["lr"])
    x = (torch.arange(n) if nodeid else torch.ones(n, dtype=torch.int64)).reshape(n, 1, 1)
    loss_fn = O.loss_fn_for(False)
    for i in range(steps):
        loss = m.step(opt, x, ei, ew, torch.from_numpy(d["pos"][i]), torch.from_numpy(d["y"][i]), loss_fn)
        assert abs(loss - d["losses"][i]) <= tol * max(1.0, abs(d["losses"][i])), (i, loss, d["losses"][i])


def test_oracle_replays_reference_density_step0():
    """Seed-0 first step of the unmodified reference on config/density.yml (--use_one).

    Only step 0 is comparable: with --use_one every node has the same input row, emb_gn
    (mean_scale = 1) outputs pure rounding noise, and Adam turns the sign of noise-level
    gradients into full-size updates, so trajectories of ANY two implementations (even two
    CPU summation orders) decorrelate from step 1 on.  DESIGN.md "parity limits"."""
    _replay("trajectory_density.npz", False, 1, 1e-5)


def test_oracle_replays_reference_density_nodeid_trajectory():
    """40 Adam steps of the unmodified reference, density graph, node-id embeddings (non-degenerate)."""
    _replay("trajectory_density_nodeid.npz", True, 40, 2e-4)
