"""GPU: model-level parity of the drop-in modules against golden outputs of the unmodified reference
(tests/golden/model_*.npz, trajectory_*.npz) and against the oracle, plus full-size property tests."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import glass_oracle as O
from tests.helpers import GOLDEN, MODEL_CASES, build_product_model, keep_masks_for, load_model_case, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-4  # BASELINE.json north_star: embeddings and logits within 1e-4 relative (fp32)


@pytest.fixture(scope="module", autouse=True)
def _lib():
    from glass_b200 import build
    build.build()
    torch.cuda.set_device(0)


def _product_from_case(c, dropout=0.0):
    m = build_product_model(c["raw"], c["x"].shape[0], dropout=dropout)
    m.load_state_dict(c["sd"])
    return m.to(DEV)


def _grad_tol(k, c):
    """--use_one: every node has the same input row, so emb_gn sees zero variance and its output (and the
    gradients of input_emb / emb_gn) are rounding noise amplified by 1/sqrt(eps) in ANY implementation
    (DESIGN.md "parity limits"); they are compared loosely.  Everything else: the 1e-4 bar."""
    if c["raw"]["emb"] == "one" and ("input_emb" in k or "emb_gn" in k):
        return 2e-2
    return TOL


def _dev(c):
    z = c["z"].to(DEV) if c["z"] is not None else None
    return c["x"].to(DEV), c["ei"].to(DEV), c["ew"].to(DEV), c["pos"].to(DEV), z


@pytest.mark.parametrize("name", MODEL_CASES)
def test_forward_matches_reference_golden(name):
    c = load_model_case(name)
    m = _product_from_case(c).eval()
    x, ei, ew, pos, z = _dev(c)
    with torch.no_grad():
        emb = m.NodeEmb(x, ei, ew, z)
        pooled = m.Pool(emb, pos, m.pools[0])
        logits = m(x, ei, ew, pos, z)
    assert rel_err(emb.cpu(), c["emb"]) < TOL
    assert rel_err(pooled.cpu(), c["pooled"]) < TOL
    assert rel_err(logits.cpu(), c["logits"]) < TOL


@pytest.mark.parametrize("name", MODEL_CASES)
def test_gradients_match_reference_golden(name):
    c = load_model_case(name)
    m = _product_from_case(c).train()
    x, ei, ew, pos, z = _dev(c)
    logits = m(x, ei, ew, pos, z)
    loss = O.loss_fn_for(c["cfg"].out_dim == 1)(logits, c["y"].to(DEV))
    loss.backward()
    assert abs(float(loss) - c["loss"]) < TOL * max(1.0, abs(c["loss"]))
    for k, p in m.named_parameters():
        tol = _grad_tol(k, c)
        assert rel_err(p.grad.cpu(), c["grads"][k]) < tol, k


@pytest.mark.parametrize("name", ["ppibp_like", "emuser_like", "coreness_like"])
def test_train_mode_with_injected_dropout_masks_matches_oracle(name):
    from glass_b200 import ops
    c = load_model_case(name)
    p = 0.5
    n = c["x"].shape[0]
    keeps = keep_masks_for(c["raw"], n, p, seed=3)
    cfg = c["cfg"]
    cfg.dropout = p
    sd = {k: v.clone().requires_grad_(True) for k, v in c["sd"].items()}
    adj = O.build_adj(c["ei"], c["ew"], n, cfg.aggr)
    ref_logits, _, _ = O.glass_forward(sd, c["x"], adj, c["pos"], c["z"], cfg, training=True, keeps=keeps)
    ref_loss = O.loss_fn_for(cfg.out_dim == 1)(ref_logits, c["y"])
    ref_loss.backward()
    m = _product_from_case(c, dropout=p).train()
    x, ei, ew, pos, z = _dev(c)
    with ops.inject_keep_masks([k.to(DEV) for k in keeps]):
        logits = m(x, ei, ew, pos, z)
    loss = O.loss_fn_for(cfg.out_dim == 1)(logits, c["y"].to(DEV))
    loss.backward()
    assert rel_err(logits.detach().cpu(), ref_logits.detach()) < TOL
    for k, prm in m.named_parameters():
        assert rel_err(prm.grad.cpu(), sd[k].grad) < _grad_tol(k, c), k


def test_generic_pool_path_and_poolmodule_api():
    """GLASS.Pool with a trans_fn (no fused path) goes pad2batch -> gather -> PoolModule.forward."""
    import torch.nn as nn

    from glass_b200 import models
    c = load_model_case("ppibp_like")
    m = _product_from_case(c).eval()
    x, ei, ew, pos, z = _dev(c)
    with torch.no_grad():
        emb = m.NodeEmb(x, ei, ew, z)
        fused = m.Pool(emb, pos, models.AddPool())
        generic = m.Pool(emb, pos, models.AddPool(trans_fn=nn.Identity()))
        size_generic = m.Pool(emb, pos, models.SizePool(trans_fn=nn.Identity()))
        size_fused = m.Pool(emb, pos, models.SizePool())
    assert rel_err(generic.cpu(), fused.cpu()) < 1e-6
    assert rel_err(size_generic.cpu(), size_fused.cpu()) < 1e-6
    assert rel_err(fused.cpu(), c["pooled"]) < TOL


def test_z_none_means_all_labelled():
    c = load_model_case("maxpool_relu")       # golden generated with z=None
    assert c["z"] is None
    m = _product_from_case(c).eval()
    x, ei, ew, pos, _ = _dev(c)
    with torch.no_grad():
        a = m(x, ei, ew, pos, None)
        b = m(x, ei, ew, pos, torch.ones(x.shape[0], dtype=torch.int64, device=DEV))
    assert torch.equal(a, b)


def test_replays_reference_density_nodeid_trajectory():
    """40 Adam steps of the unmodified reference (density graph, node-id embeddings, config/density.yml)
    replayed through glass_b200.train-style steps on the GPU: losses within 1e-3 relative."""
    from glass_b200 import datasets, utils
    d = np.load(os.path.join(GOLDEN, "trajectory_density_nodeid.npz"))
    params = json.loads(str(d["params"]))
    ei, ew, n = datasets.load_edges("density")
    raw = dict(H=params["hidden_dim"], L=params["conv_layer"], aggr=params["aggr"], z=params["z_ratio"], act="elu",
               jk=1, out=3, emb="nodeid", pool=params["pool"])
    m = build_product_model(raw, n, dropout=params["dropout"])
    m.load_state_dict({k[3:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("sd.")})
    m = m.to(DEV).train()
    opt = torch.optim.Adam(m.parameters(), lr=params["lr"])
    x = torch.arange(n, device=DEV).reshape(n, 1, 1)
    ei, ew = ei.to(DEV), ew.to(DEV)
    loss_fn = torch.nn.CrossEntropyLoss()
    for i in range(40):
        pos = torch.from_numpy(d["pos"][i]).to(DEV)
        y = torch.from_numpy(d["y"][i]).to(DEV)
        opt.zero_grad()
        loss = loss_fn(m(x, ei, ew, pos, utils.MaxZOZ(x, pos), id=0), y)
        loss.backward()
        opt.step()
        ref = float(d["losses"][i])
        assert abs(float(loss) - ref) <= 1e-3 * max(1.0, abs(ref)), (i, float(loss), ref)


def test_reference_density_step0_loss():
    """--use_one, step 0 of config/density.yml.  All nodes share one embedding row, so emb_gn's input has
    zero variance: the exact result is `bias`, but the reference's fp32 sequential mean is off by ~1e-5
    relative and GraphNorm multiplies that residue by 1/sqrt(eps) = 316, which moves ITS loss by ~0.3 %
    (the CPU oracle repeats the same summation and reproduces it; tests/test_oracle_golden.py).  The CUDA
    path accumulates the mean in fp64 and returns the exact value, hence the 1e-2 bar here; non-degenerate
    inputs are held to 1e-3 over 40 optimizer steps in test_replays_reference_density_nodeid_trajectory."""
    d = np.load(os.path.join(GOLDEN, "trajectory_density.npz"))
    params = json.loads(str(d["params"]))
    from glass_b200 import datasets, utils
    ei, ew, n = datasets.load_edges("density")
    raw = dict(H=params["hidden_dim"], L=params["conv_layer"], aggr=params["aggr"], z=params["z_ratio"], act="elu",
               jk=1, out=3, emb="one", pool=params["pool"])
    m = build_product_model(raw, n)
    m.load_state_dict({k[3:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("sd.")})
    m = m.to(DEV).train()
    x = torch.ones((n, 1, 1), dtype=torch.int64, device=DEV)
    pos = torch.from_numpy(d["pos"][0]).to(DEV)
    loss = torch.nn.CrossEntropyLoss()(m(x, ei.to(DEV), ew.to(DEV), pos, utils.MaxZOZ(x, pos)),
                                       torch.from_numpy(d["y"][0]).to(DEV))
    assert abs(float(loss) - float(d["losses"][0])) < 1e-2 * float(d["losses"][0])


# ------------------------------------------------------------------------------------------
# BASELINE.json full sizes: size-independent properties (the oracle is too slow there)
# ------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def em_user_graph():
    from glass_b200 import datasets
    g = datasets.load_dataset("em_user_shaped")
    return g


@pytest.mark.parametrize("aggr", ["mean", "gcn"])
def test_full_size_spmm_properties(em_user_graph, aggr):
    from glass_b200 import ops
    g = em_user_graph
    n = g.num_nodes
    ei, ew = g.edge_index.to(DEV), g.edge_attr.to(DEV)
    assert n == 57333 and ei.shape[1] == 2 * 4573417
    adj = ops.build_csr(ei, ew, n, aggr)
    # CSR structure: sorted, duplicate free input -> identical indices; row pointers monotone
    assert torch.equal(adj.col.long(), ei[1]) and int(adj.rowptr[-1]) == ei.shape[1]
    deg = torch.bincount(ei[0], minlength=n).float()
    assert torch.equal(adj.deg, torch.where(deg < 0.5, deg + 1, deg))
    gen = torch.Generator(device=DEV).manual_seed(0)
    x = torch.randn(n, 64, device=DEV, generator=gen)
    y = torch.randn(n, 64, device=DEV, generator=gen)
    ax = ops.spmm(adj, x)
    # linearity
    assert rel_err(ops.spmm(adj, 2 * x + y).cpu(), (2 * ax + ops.spmm(adj, y)).cpu()) < 1e-5
    # adjointness of the transposed copy: <A x, y> == <x, A^T y>
    aty = ops.spmm(adj.t(), y)
    lhs, rhs = (ax.double() * y.double()).sum(), (x.double() * aty.double()).sum()
    assert abs(float(lhs - rhs)) < 1e-6 * float(ax.double().norm() * y.double().norm())
    # row sums: mean-normalised rows sum to 1
    ones = ops.spmm(adj, torch.ones(n, 4, device=DEV))
    if aggr == "mean":
        assert rel_err(ones[deg > 0].cpu(), torch.ones_like(ones[deg > 0]).cpu()) < 1e-5
    # against fp64 torch on a sample of rows
    rows = torch.randint(0, n, (64,), generator=torch.Generator().manual_seed(1))
    rp, col, val = adj.rowptr.cpu(), adj.col.cpu().long(), adj.val.cpu().double()
    xc = x.cpu().double()
    for r in rows.tolist():
        s, e = int(rp[r]), int(rp[r + 1])
        ref = (val[s:e, None] * xc[col[s:e]]).sum(0)
        assert rel_err(ax[r].cpu(), ref) < 1e-5


def test_full_size_train_step_runs_and_is_finite(em_user_graph):
    """One em_user-shaped train step (config/em_user.yml hyper-parameters) end to end."""
    import functools

    import torch.nn as nn

    from glass_b200 import datasets, models, utils
    g = em_user_graph
    n = g.num_nodes
    x = torch.arange(n, device=DEV).reshape(n, 1, 1)
    ei, ew = g.edge_index.to(DEV), g.edge_attr.to(DEV)
    conv = models.EmbZGConv(64, 64, 1, max_deg=n - 1, activation=nn.ELU(inplace=True), jk=1, dropout=0.5,
                            conv=functools.partial(models.GLASSConv, aggr="gcn", z_ratio=0.75, dropout=0.5), gn=True)
    conv.input_emb = nn.Embedding.from_pretrained(datasets.synthetic_embedding(n, 64), freeze=False)
    m = models.GLASS(conv, nn.ModuleList([nn.Linear(64, 1)]), nn.ModuleList([models.SizePool()])).to(DEV).train()
    pos = g.pos[:6].to(DEV)
    y = g.y[:6].to(DEV)
    z = utils.MaxZOZ(x, pos)
    assert int(z.sum()) == int(torch.unique(pos[pos >= 0]).numel())
    loss = nn.BCEWithLogitsLoss()(m(x, ei, ew, pos, z).flatten(), y.flatten())
    loss.backward()
    assert torch.isfinite(loss)
    for k, p in m.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), k


def test_cuda_graph_step_matches_eager_steps():
    """The captured whole-step graph (labels, fwd, loss, bwd, Adam) reproduces eager training."""
    from glass_b200 import utils
    from glass_b200.graphed import GraphedTrainStep
    c = load_model_case("ppibp_like")
    x, ei, ew, pos, _ = _dev(c)
    y = c["y"].to(DEV)
    loss_fn = torch.nn.CrossEntropyLoss()
    g = torch.Generator().manual_seed(0)
    batches = []
    for _ in range(6):
        p = c["pos"].clone()
        perm = torch.randperm(c["x"].shape[0], generator=g)
        p[p >= 0] = perm[p[p >= 0]]
        batches.append((p.to(DEV), y))
    # eager run
    m1 = _product_from_case(c).train()
    opt = torch.optim.Adam(m1.parameters(), lr=1e-2)
    eager = []
    for p, t in batches:
        opt.zero_grad()
        loss = loss_fn(m1(x, ei, ew, p, utils.MaxZOZ(x, p), id=0), t)
        loss.backward()
        opt.step()
        eager.append(float(loss))
    # graphed run from the same initial state
    m2 = _product_from_case(c).train()
    init = {k: v.clone() for k, v in m2.state_dict().items()}
    step = GraphedTrainStep(m2, loss_fn, x, ei, ew, batches[0][0], y, lr=1e-2)
    step.capture()
    step.reset_to(init)
    graphed = [float(step(p.cpu().pin_memory(), t)) for p, t in batches]
    for a, b in zip(eager, graphed):
        assert abs(a - b) <= 1e-4 * max(1.0, abs(a)), (eager, graphed)
    for k, v in m1.state_dict().items():
        assert rel_err(m2.state_dict()[k].cpu(), v.cpu()) < 1e-3, k


def test_cuda_graph_replays_draw_fresh_dropout_masks():
    """torch's graph-safe Philox state advances on every replay: two replays of the same batch with
    dropout 0.5 must see different keep masks (different losses), and lr changes must take effect."""
    from glass_b200.graphed import GraphedForward, GraphedTrainStep
    c = load_model_case("emuser_like")
    x, ei, ew, pos, _ = _dev(c)
    y = c["y"].to(DEV)
    m = _product_from_case(c, dropout=0.5).train()
    step = GraphedTrainStep(m, O.loss_fn_for(True), x, ei, ew, pos, y, lr=0.0).capture()   # lr 0: weights frozen
    losses = [float(step(pos, y)) for _ in range(4)]
    assert len({round(v, 6) for v in losses}) > 1, losses
    before = {k: v.clone() for k, v in m.state_dict().items()}
    step.set_lr(1e-2)
    step(pos, y)
    assert any(not torch.equal(before[k], v) for k, v in m.state_dict().items())
    # graphed inference equals eager inference
    m.eval()
    with torch.no_grad():
        ref = m(x, ei, ew, pos, None)
    fwd = GraphedForward(m, x, ei, ew, pos, z_fn=lambda a, b: None)
    assert torch.allclose(fwd(pos), ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", ["ppibp_like", "emuser_like"])
def test_graphed_eval_epoch_matches_train_test_on_ragged_loader(name):
    """graphed.test_epoch (captured forward, last batch padded with empty subgraphs, one gather per epoch)
    returns what train.test returns on the same loader in the same seeded order."""
    from glass_b200 import SubGDataset, train, utils
    from glass_b200.graphed import GraphedForward, test_epoch
    c = load_model_case(name)
    x, ei, ew, pos, _ = _dev(c)
    y = c["y"].to(DEV)
    reps = 7 // pos.shape[0] + 1                           # 7+ subgraphs, batch size 3 -> ragged last batch
    pos7, y7 = pos.repeat(reps, 1)[:7].clone(), y.repeat(*([reps] + [1] * (y.dim() - 1)))[:7].clone()
    pos7[1:] = pos7[1:].roll(1, dims=1)
    ds = SubGDataset.GDataset(x, ei, ew, pos7, y7)
    loader = SubGDataset.ZGDataloader(ds, 3, z_fn=utils.MaxZOZ, shuffle=True, drop_last=False)
    m = _product_from_case(c).eval()
    loss_fn = O.loss_fn_for(c["cfg"].out_dim == 1)
    metric = lambda pred, t: float(np.abs(pred).sum())
    torch.manual_seed(5)
    ref_score, ref_loss = train.test(m, loader, metric, loss_fn)
    fwd = GraphedForward(m, x, ei, ew, pos7[:3])
    torch.manual_seed(5)
    score, loss = test_epoch(fwd, loader, metric, loss_fn)
    assert abs(score - ref_score) <= 1e-5 * max(1.0, abs(ref_score))
    assert abs(float(loss) - float(ref_loss)) <= 1e-6 * max(1.0, abs(float(ref_loss)))


def test_edge_cases_single_subgraph_empty_rows_and_tiny_graph():
    """B = 1, a one-node subgraph, an all-padding row in the middle, and a 3-node graph."""
    import functools

    import torch.nn as nn

    from glass_b200 import models, utils
    n = 3
    ei = torch.tensor([[0, 1, 1, 2], [1, 0, 2, 1]], device=DEV)
    ew = torch.ones(4, device=DEV)
    x = torch.arange(n, device=DEV).reshape(n, 1, 1)
    conv = models.EmbZGConv(8, 8, 2, max_deg=n - 1, activation=nn.ELU(inplace=True), jk=1, dropout=0.0,
                            conv=functools.partial(models.GLASSConv, aggr="mean", z_ratio=0.8, dropout=0.0), gn=True)
    m = models.GLASS(conv, nn.ModuleList([nn.Linear(16, 2)]), nn.ModuleList([models.MeanPool()])).to(DEV).train()
    for pos in (torch.tensor([[2]], device=DEV), torch.tensor([[0, 1, -1], [-1, -1, -1], [2, -1, -1]], device=DEV)):
        z = utils.MaxZOZ(x, pos)
        out = m(x, ei, ew, pos, z)
        assert out.shape == (pos.shape[0], 2) and torch.isfinite(out).all()
        out.sum().backward()
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
    # oracle agrees on the ragged batch
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    cfg = O.GlassConfig(hidden_dim=8, conv_layer=2, aggr="mean", z_ratio=0.8, pool="mean", jk=True, out_dim=2)
    posc = torch.tensor([[0, 1, -1], [2, -1, -1]])
    ref, _, _ = O.glass_forward(sd, x.cpu(), O.build_adj(ei.cpu(), ew.cpu(), n, "mean"), posc, O.max_zero_one(n, posc), cfg)
    got = m.eval()(x, ei, ew, posc.to(DEV), utils.MaxZOZ(x, posc.to(DEV)))
    assert rel_err(got.detach().cpu(), ref) < 1e-4


def test_pretraining_modules_match_reference_golden():
    """MyGCNConv / EmbGConv / EdgeGNN (impl/models.py:361-509) vs the reference."""
    import functools

    import torch.nn as nn

    from glass_b200 import models
    d = np.load(os.path.join(GOLDEN, "model_edgegnn.npz"))
    H, L = 32, 2
    conv = models.EmbGConv(H, H, H, L, max_deg=11, activation=nn.ReLU(inplace=True), jk=True, dropout=0.0,
                           conv=functools.partial(models.MyGCNConv, aggr="mean", activation=nn.ReLU(inplace=True)), gn=True)
    m = models.EdgeGNN(conv, nn.ModuleList([nn.Linear(H * L, 1)]), nn.ModuleList([models.MeanPool()]))
    sd = {k[3:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("sd.")}
    assert list(m.state_dict().keys()) == list(sd.keys())
    m.load_state_dict(sd)
    m = m.to(DEV).train()
    t = lambda k: torch.from_numpy(d[k]).to(DEV)
    logits = m(t("x"), t("ei"), t("ew"), t("pairs"))
    loss = nn.BCEWithLogitsLoss()(logits.flatten(), t("y"))
    loss.backward()
    assert rel_err(logits.detach().cpu(), d["logits"]) < TOL
    assert abs(float(loss) - float(d["loss"])) < TOL
    errs = {k: rel_err(p.grad.cpu(), d[f"grad.{k}"]) for k, p in m.named_parameters()}
    assert max(errs.values()) < TOL, {k: v for k, v in errs.items() if v >= TOL}


# ------------------------------------------------------------------------------------------
# multi-label-batch evaluation: shared base + sparse label correction
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["emuser_like", "ppibp_like", "coreness_like", "density_like", "cutratio_like"])
def test_shared_base_forward_equals_per_batch_forward(name):
    """adj @ U once + adj[:, labelled] @ delta[labelled] per batch gives the logits of the ordinary forward
    (and of the reference golden) for several different label batches of the same weights."""
    from glass_b200 import utils
    c = load_model_case(name)
    m = _product_from_case(c).eval()
    x, ei, ew, pos, z = _dev(c)
    n = x.shape[0]
    with torch.no_grad():
        base = m.shared_base(x, ei, ew)
        got = m.forward_from_base(base, ei, ew, pos, z)
        assert rel_err(got.cpu(), c["logits"]) < TOL                    # the unmodified reference's logits
        g = torch.Generator().manual_seed(4)
        for _ in range(3):                                                # other label batches, same base
            p = c["pos"].clone()
            perm = torch.randperm(n, generator=g)
            p[p >= 0] = perm[p[p >= 0]]
            p = p.to(DEV)
            zz = utils.MaxZOZ(x, p)
            ref = m(x, ei, ew, p, zz)
            got = m.forward_from_base(base, ei, ew, p, zz)
            assert rel_err(got.cpu(), ref.cpu()) < 1e-5
        # z = None: every node labelled (the correction touches every column)
        assert rel_err(m.forward_from_base(base, ei, ew, pos, None).cpu(), m(x, ei, ew, pos, None).cpu()) < 5e-5


def test_shared_base_eval_epoch_matches_graphed_eval_epoch():
    """graphed.test_epoch with the shared-base evaluator returns what it returns with the per-batch evaluator; after
    the weights change, refresh() (called by test_epoch) picks the new ones up."""
    from glass_b200 import SubGDataset, utils
    from glass_b200.graphed import GraphedForward, GraphedSharedBaseForward, test_epoch
    c = load_model_case("emuser_like")
    x, ei, ew, pos, _ = _dev(c)
    y = c["y"].to(DEV)
    reps = 11 // pos.shape[0] + 1
    pos_all = pos.repeat(reps, 1)[:11].clone()
    y_all = y.repeat(*([reps] + [1] * (y.dim() - 1)))[:11].clone()
    for i in range(1, 11):
        pos_all[i] = pos_all[i].roll(i, dims=0)
    ds = SubGDataset.GDataset(x, ei, ew, pos_all, y_all)
    loader = SubGDataset.ZGDataloader(ds, 4, z_fn=utils.MaxZOZ, shuffle=True, drop_last=False)
    m = _product_from_case(c).eval()
    loss_fn = O.loss_fn_for(c["cfg"].out_dim == 1)
    metric = lambda pred, t: float(np.abs(pred).sum())
    plain = GraphedForward(m, x, ei, ew, pos_all[:4])
    shared = GraphedSharedBaseForward(m, x, ei, ew, pos_all[:4])
    for round_ in range(2):
        torch.manual_seed(5)
        ref_score, ref_loss = test_epoch(plain, loader, metric, loss_fn)
        torch.manual_seed(5)
        score, loss = test_epoch(shared, loader, metric, loss_fn)
        assert abs(score - ref_score) <= 1e-5 * max(1.0, abs(ref_score))
        assert abs(float(loss) - float(ref_loss)) <= 1e-5 * max(1.0, abs(float(ref_loss)))
        with torch.no_grad():                                             # "one optimizer step later"
            for prm in m.parameters():
                prm.add_(0.05 * torch.randn_like(prm))


def test_shared_base_unsupported_width_raises():
    c = load_model_case("component_like")           # H = 17
    m = _product_from_case(c).eval()
    x, ei, ew, pos, z = _dev(c)
    with pytest.raises(NotImplementedError):
        m.shared_base(x, ei, ew)


@pytest.mark.parametrize("name", ["cutratio_like", "ppibp_like", "emuser_like", "component_like"])
def test_seeded_training_is_bit_reproducible(name):
    """SURVEY.md section 5 "Determinism": two runs from the same seed give bit-identical losses and parameters.
    Pooling / embedding backward add in a fixed order (no float atomics), statistics are reduced in a fixed order,
    the dropout generator is counter based."""
    from glass_b200 import ops, utils
    c = load_model_case(name)
    x, ei, ew, pos, _ = _dev(c)
    y = c["y"].to(DEV)
    loss_fn = O.loss_fn_for(c["cfg"].out_dim == 1)
    n = x.shape[0]
    g = torch.Generator().manual_seed(11)
    batches = []
    for _ in range(4):
        p = c["pos"].clone()
        perm = torch.randperm(n, generator=g)
        p[p >= 0] = perm[p[p >= 0]]
        p[1] = p[0]                                   # the same nodes in two subgraphs: several gradients per node
        batches.append(p.to(DEV))
    runs = []
    for _ in range(2):
        m = _product_from_case(c, dropout=0.3).train()
        opt = torch.optim.Adam(m.parameters(), lr=1e-2)
        ops.manual_seed(7)
        losses = []
        for p in batches:
            opt.zero_grad()
            loss = loss_fn(m(x, ei, ew, p, utils.MaxZOZ(x, p), id=0), y)
            loss.backward()
            opt.step()
            losses.append(float(loss))
        runs.append((losses, {k: v.clone() for k, v in m.state_dict().items()}))
    assert runs[0][0] == runs[1][0]
    for k, v in runs[0][1].items():
        assert torch.equal(v, runs[1][1][k]), k
