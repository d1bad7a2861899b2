"""CPU: host-side logic of the drop-in (no kernels run): ABI surface, parameter/seed parity, loaders."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

from tests.helpers import GOLDEN, MODEL_CASES, build_product_model, load_model_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "glass_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(glass_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from glass_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/glass_b200.h but not exported"
    assert sorted(_lib.PROTOTYPES) == names, "ctypes prototypes out of sync with the header"
    assert _lib.load().glass_abi_version() == 2
    # size queries are pure host arithmetic and must work without a GPU
    assert _lib.load().glass_pair_linear_mix_bwd_workspace_bytes(57333, 64, 128) > 0
    assert _lib.load().glass_graphnorm_workspace_bytes(57333, 64) > 0


def test_no_cpu_fallback():
    from glass_b200 import models, ops
    with pytest.raises(RuntimeError, match="no CPU"):
        ops.spmm(ops.CSRAdj(2, *([torch.zeros(3, dtype=torch.int32)] * 6), None, "sum"), torch.zeros(2, 4))
    with pytest.raises(RuntimeError):
        ops.maxzoz(4, torch.zeros(2, 2, dtype=torch.int64))
    with pytest.raises(NotImplementedError):
        models.buildAdj(torch.zeros(2, 1, dtype=torch.int64), torch.ones(1), 3, "max")
    from glass_b200 import config
    with pytest.raises(RuntimeError):
        config.set_device(-1)


def test_product_does_not_import_oracle_or_sparse_fallbacks():
    pkg = os.path.join(ROOT, "glass_b200")
    for f in os.listdir(pkg):
        if not f.endswith(".py"):
            continue
        for line in open(os.path.join(pkg, f)):
            code = line.split("#")[0]
            assert not re.match(r"\s*(from|import)\s+oracle", code), (f, line)
            assert "import_module(\"oracle" not in code, (f, line)
            for banned in ("torch.sparse_coo_tensor(", "torch.sparse.", "torch_geometric", "triton"):
                if banned in code and not code.lstrip().startswith(('"', "'")) and '"""' not in code:
                    in_doc = any(w in code for w in ("fallback", "PyG 1.7.2", "stand-in"))
                    assert in_doc, (f, line)


@pytest.mark.parametrize("name", MODEL_CASES)
def test_seeded_init_and_state_dict_keys_match_reference(name):
    """Same seed + same construction order => bit-identical parameters and identical state_dict keys
    as the unmodified reference (golden state_dicts were produced by GLASSTest.buildModel's recipe)."""
    c = load_model_case(name)
    torch.manual_seed(1234)
    m = build_product_model(c["raw"], c["x"].shape[0])
    with torch.no_grad():
        for k, p in m.named_parameters():
            if "gn" in k:
                p.add_(0.3 * torch.randn_like(p))
    sd = m.state_dict()
    assert list(sd.keys()) == list(c["sd"].keys())
    for k in sd:
        assert torch.equal(sd[k], c["sd"][k]), k


def test_shipped_dataset_split_matches_reference_seed():
    from glass_b200 import datasets
    d = np.load(os.path.join(GOLDEN, "trajectory_density.npz"))
    torch.manual_seed(0)
    g = datasets.load_dataset("density")
    assert np.array_equal(g.mask.numpy(), d["mask"])
    assert g.edge_index.shape == (2, 59924) and g.pos.shape == (250, 20)


def test_loader_yields_reference_batches():
    """ZGDataloader draws the same permutation as the reference's (golden batches of trajectory_density)."""
    from glass_b200 import SubGDataset, datasets
    d = np.load(os.path.join(GOLDEN, "trajectory_density.npz"))
    params = json.loads(str(d["params"]))
    torch.manual_seed(0)
    g = datasets.load_dataset("density")
    g.y = g.y.to(torch.int64)
    g.setOneFeature()
    trn = SubGDataset.GDataset(*g.get_split("train"))
    # the reference builds the model between split and loader: burn the same RNG draws
    raw = dict(H=params["hidden_dim"], L=params["conv_layer"], aggr=params["aggr"], z=params["z_ratio"], act="elu",
               jk=1, out=3, emb="one", pool=params["pool"])
    m = build_product_model(raw, g.x.shape[0])
    for k, v in m.state_dict().items():
        assert torch.equal(v, torch.from_numpy(d[f"sd.{k}"])), k
    seen = []
    loader = SubGDataset.ZGDataloader(trn, params["batch_size"], z_fn=lambda x, p: torch.zeros(1), shuffle=True,
                                      drop_last=True)
    for batch in loader:
        seen.append(batch[3].numpy())
    assert np.array_equal(np.stack(seen), d["pos"])


def test_synthetic_shapes_small():
    from glass_b200 import datasets
    e = datasets.uniform_edges(1000, 5000, 0)
    assert e.shape == (2, 5000) and int((e[0] == e[1]).sum()) == 0
    key = e[0] * 1000 + e[1]
    assert torch.unique(key).numel() == 5000
    e = datasets.powerlaw_edges(2000, 8000, 0)
    deg = torch.bincount(torch.cat((e[0], e[1])), minlength=2000)
    assert e.shape == (2, 8000) and deg.max() > 20 * deg.float().mean()


def test_loaders_follow_the_torch_dataloader_order():
    """GDataloader / epoch_batches draw from torch's RNG exactly like the torch.utils.data.DataLoader the
    reference subclasses (impl/SubGDataset.py:38-47): same index batches, same RNG state afterwards."""
    from torch.utils.data import DataLoader
    from glass_b200 import SubGDataset
    n = 23
    ds = SubGDataset.GDataset(torch.zeros(n * 4, 1), torch.zeros(2, 0, dtype=torch.int64), torch.zeros(0),
                              torch.arange(n * 4).reshape(n, 4), torch.arange(n))
    for shuffle, drop_last in ((True, True), (True, False), (False, False)):
        ref_loader = DataLoader(torch.arange(n), batch_size=5, shuffle=shuffle, drop_last=drop_last)
        loader = SubGDataset.ZGDataloader(ds, batch_size=5, shuffle=shuffle, drop_last=drop_last)
        assert len(loader) == len(ref_loader)
        torch.manual_seed(11)
        ref = [(ds.pos[idx], ds.y[idx]) for _ in range(2) for idx in ref_loader]            # two epochs
        after_ref = torch.rand(1)
        torch.manual_seed(11)
        via_iter = [(b[3], b[-1]) for _ in range(2) for b in loader]
        after_iter = torch.rand(1)
        torch.manual_seed(11)
        via_epoch = [(p, y) for _ in range(2) for p, y in SubGDataset.epoch_batches(loader)]
        after_epoch = torch.rand(1)
        assert len(ref) == len(via_iter) == len(via_epoch) == 2 * (n // 5 if drop_last else -(-n // 5))
        for (rp, ry), (ip, iy), (ep, ey) in zip(ref, via_iter, via_epoch):
            assert torch.equal(rp, ip) and torch.equal(ry, iy) and torch.equal(rp, ep) and torch.equal(ry, ey)
        assert torch.equal(after_ref, after_iter) and torch.equal(after_ref, after_epoch)   # same RNG consumption
    b = next(iter(SubGDataset.ZGDataloader(ds, 4, shuffle=False)))
    assert len(b) == 6 and b[4].shape == (n * 4, 1) and b[4].dtype == torch.int64 and loader.get_pos() is ds.pos


def test_train_and_test_epoch_drivers_follow_the_reference_loop():
    """glass_b200.train.{train,test} against a hand-written loop on a plain torch module (host logic only)."""
    from glass_b200 import train

    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = torch.nn.Linear(3, 2)

        def forward(self, a, b, id=0):
            return self.lin(a + b)

    g = torch.Generator().manual_seed(0)
    batches = [(torch.randn(n, 3, generator=g), torch.randn(n, 3, generator=g), torch.randint(0, 2, (n,), generator=g))
               for n in (4, 4, 3)]
    loss_fn = torch.nn.CrossEntropyLoss()
    results = []
    for sync in (True, False, None):
        torch.manual_seed(1)
        m = Tiny()
        opt = torch.optim.Adam(m.parameters(), lr=1e-2)
        if sync is None:                                   # the loop of impl/train.py:8-17, written out
            m.train()
            losses = []
            for a, b, y in batches:
                opt.zero_grad()
                loss = loss_fn(m(a, b, id=0), y)
                loss.backward()
                losses.append(loss.item())
                opt.step()
            mean = sum(losses) / len(losses)
        else:
            mean = train.train(opt, m, batches, loss_fn, sync_each_step=sync)
        score, loss = train.test(m, batches, lambda p, t: float((p.argmax(-1) == t).mean()), loss_fn)
        results.append((mean, score, float(loss)))
    for r in results[:2]:
        assert abs(r[0] - results[2][0]) < 1e-6 and r[1] == results[2][1] and abs(r[2] - results[2][2]) < 1e-6


def test_pretraining_modules_state_dict_layout_matches_reference_golden():
    """EmbGConv / MyGCNConv / EdgeGNN are built with the reference's parameter names and order (CPU part of
    tests/test_gpu_model.py::test_pretraining_modules_match_reference_golden)."""
    import functools

    import numpy as np
    import torch.nn as nn

    from glass_b200 import models
    from tests.helpers import GOLDEN
    d = np.load(os.path.join(GOLDEN, "model_edgegnn.npz"))
    H, L = 32, 2
    conv = models.EmbGConv(H, H, H, L, max_deg=11, activation=nn.ReLU(inplace=True), jk=True, dropout=0.0,
                           conv=functools.partial(models.MyGCNConv, aggr="mean", activation=nn.ReLU(inplace=True)), gn=True)
    m = models.EdgeGNN(conv, nn.ModuleList([nn.Linear(H * L, 1)]), nn.ModuleList([models.MeanPool()]))
    sd = {k[3:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("sd.")}
    assert list(m.state_dict().keys()) == list(sd.keys())
    m.load_state_dict(sd)
    # pool wrappers keep the reference's class relations (impl/models.py:295-319)
    assert issubclass(models.SizePool, models.AddPool) and issubclass(models.MaxPool, models.PoolModule)
    assert models.MeanPool().padded_mode() == "mean" and models.SizePool().padded_mode() == "size"
    assert models.AddPool(trans_fn=nn.Identity()).padded_mode() is None
    one = models.EmbGConv(8, 16, 4, 1, max_deg=3)
    assert [tuple(c.trans_fn.weight.shape) for c in one.convs] == [(4, 8)] and len(one.gns) == 0


def test_binaryf1_and_microf1_match_sklearn():
    """impl/metrics.py:5-20 call sklearn f1_score(average="micro"): a single-column binary target is scored as
    a binary problem over both classes (= accuracy), a wider indicator matrix as multi-label tp/fp/fn."""
    from sklearn.metrics import f1_score
    from glass_b200 import metrics
    g = np.random.default_rng(0)
    n = 200
    logits = g.normal(size=(n, 1))
    for label in (g.integers(0, 2, n).astype(np.float32), g.integers(0, 2, (n, 1)).astype(np.float32)):
        ref = f1_score(label.reshape(n, -1), (logits > 0).astype(np.int64), average="micro")
        assert abs(metrics.binaryf1(logits, label) - ref) < 1e-12
    ml_logits, ml_label = g.normal(size=(n, 5)), g.integers(0, 2, (n, 5)).astype(np.float32)
    ref = f1_score(ml_label, (ml_logits > 0).astype(np.int64), average="micro")
    assert abs(metrics.binaryf1(ml_logits, ml_label) - ref) < 1e-12
    mc_logits, mc_label = g.normal(size=(n, 6)), g.integers(0, 6, n)
    ref = f1_score(mc_label, mc_logits.argmax(1), average="micro")
    assert abs(metrics.microf1(mc_logits, mc_label) - ref) < 1e-12


def test_gdataset_rejects_node_ids_outside_the_graph():
    from glass_b200 import SubGDataset
    with pytest.raises(IndexError):
        SubGDataset.GDataset(torch.zeros(5, 1), torch.zeros(2, 0, dtype=torch.int64), torch.zeros(0),
                             torch.tensor([[0, 7, -1]]), torch.zeros(1))


def _write_fake_real_dataset(root, name, multilabel):
    """A tiny dataset in the SubGNN text format the reference parses (datasets.py:131-222)."""
    d = os.path.join(root, "dataset", name)
    os.makedirs(d)
    g = np.random.default_rng(3)
    n = 40
    edges = {(int(a), int(b)) for a, b in g.integers(0, n, (150, 2))}
    with open(os.path.join(d, "edge_list.txt"), "w") as f:
        for a, b in sorted(edges):
            f.write(f"{a} {b}\n")
        f.write("3 1\n1 3\n")                                       # both directions of one pair
    names = ["alpha", "beta", "gamma"]
    with open(os.path.join(d, "subgraphs.pth"), "w") as f:
        for i in range(30):
            nodes = "-".join(str(int(v)) for v in g.choice(n, size=int(g.integers(2, 7)), replace=False))
            k = int(g.integers(1, 3)) if multilabel else 1
            labs = "-".join(g.choice(names, size=k, replace=False))
            split = ["train", "train", "train", "val", "test", "test"][i % 6]     # val smaller than test -> swapped
            f.write(f"{nodes}\t{labs}\t{split}\n")
    return d


@pytest.mark.parametrize("multilabel", [False, True])
def test_real_dataset_loader_parses_the_subgnn_format(tmp_path, monkeypatch, multilabel):
    from glass_b200 import datasets
    _write_fake_real_dataset(str(tmp_path), "ppi_bp", multilabel)
    monkeypatch.setenv("GLASS_DATASET_DIR", str(tmp_path / "dataset"))
    g = datasets.load_dataset("ppi_bp")
    assert g.pos.shape[0] == 30 and g.mask.tolist().count(0) == 15
    assert g.mask.tolist().count(1) == 10 and g.mask.tolist().count(2) == 5     # val/test swapped (val was smaller)
    assert g.y.dtype == torch.float32 and (g.y.ndim == 2) == multilabel
    ei = g.edge_index
    key = ei[0] * g.num_nodes + ei[1]
    assert torch.equal(key, torch.unique(key))                                   # sorted, duplicate free
    assert torch.equal(torch.unique(ei[1] * g.num_nodes + ei[0]), key)           # symmetric
    off = ei[0] != ei[1]
    assert torch.all(g.edge_attr[off] == 1.0)
    with pytest.raises(FileNotFoundError):
        datasets.load_dataset("em_user")


@pytest.mark.needs_reference
@pytest.mark.parametrize("multilabel", [False, True])
def test_real_dataset_loader_equals_the_reference_loader(tmp_path, monkeypatch, multilabel):
    import sys
    from glass_b200 import datasets
    _write_fake_real_dataset(str(tmp_path), "hpo_metab", multilabel)
    monkeypatch.setenv("GLASS_DATASET_DIR", str(tmp_path / "dataset"))
    mine = datasets.load_dataset("hpo_metab")
    ref_root = os.environ.get("GLASS_REFERENCE", "/root/reference")
    shim = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "pyg_shim")
    monkeypatch.chdir(tmp_path)
    monkeypatch.syspath_prepend(shim)
    monkeypatch.syspath_prepend(ref_root)
    for m in [k for k in sys.modules if k == "datasets" or k.startswith("torch_geometric")]:
        monkeypatch.delitem(sys.modules, m)
    import datasets as ref_datasets
    ref = ref_datasets.load_dataset("hpo_metab")
    assert torch.equal(mine.edge_index, ref.edge_index) and torch.equal(mine.edge_attr, ref.edge_attr)
    assert torch.equal(mine.pos, ref.pos) and torch.equal(mine.y, ref.y)
    assert torch.equal(mine.mask.to(ref.mask.dtype), ref.mask)
    assert mine.x.shape == ref.x.shape


# ------------------------------------------------------------------------------------------ row partition host logic
def test_partition_phase_groups_and_column_split():
    """glass_b200.partition: the pipelined product takes the own shard first and the peers in arrival order in at
    most `max_phases` groups; split_columns_by_group re-bases the columns of every group to its stage buffer and
    keeps every entry exactly once, in row order."""
    from glass_b200.partition import phase_groups, split_columns_by_group, split_columns_by_owner
    for world in (1, 2, 3, 4, 8, 16):
        for rank in (0, world - 1):
            g = phase_groups(rank, world, 4)
            assert g[0] == [rank] and len(g) <= 4
            flat = [s_ for grp in g for s_ in grp]
            assert sorted(flat) == list(range(world))
            assert flat[1:] == [(rank + i) % world for i in range(1, world)]        # arrival order
    assert phase_groups(0, 8, 4) == [[0], [1, 2, 3], [4, 5, 6], [7]]
    assert phase_groups(5, 8, 2) == [[5], [6, 7, 0, 1, 2, 3, 4]]
    # a 4-row block whose columns live in a padded layout of 3 owners x pad 4
    rp = torch.tensor([0, 3, 5, 5, 9], dtype=torch.int32)
    c = torch.tensor([0, 5, 9, 2, 3, 1, 4, 8, 11], dtype=torch.int32)
    v = torch.arange(9, dtype=torch.float32)
    groups = [[1], [2, 0]]                                     # rank 1 of 3: own, then 2 and 0 in one group
    parts = split_columns_by_group(rp, c, v, 4, groups)
    (rp0, c0, v0), (rp1, c1, v1) = parts
    assert rp0.tolist() == [0, 1, 1, 1, 2] and c0.tolist() == [1, 0] and v0.tolist() == [1.0, 6.0]
    # owner 2 -> slot 0 (columns 8..11 -> 0..3), owner 0 -> slot 1 (columns 0..3 -> 4..7)
    assert rp1.tolist() == [0, 2, 4, 4, 7] and c1.tolist() == [4, 1, 6, 7, 5, 0, 3] and v1.tolist() == [0., 2., 3., 4., 5., 7., 8.]
    by_owner = split_columns_by_owner(rp, c, v, [0, 4, 8, 12])
    assert [p[1].tolist() for p in by_owner] == [[0, 2, 3, 1], [1, 0], [1, 0, 3]]
    assert sum(int(p[0][-1]) for p in by_owner) == 9


def test_dp_owned_ranges_cover_the_table():
    """glass_b200.dp.owned_range: contiguous, multiples of 4, disjoint, covering -- every element of the table is reduced
    and updated by exactly one rank."""
    from glass_b200.dp import owned_range
    for n in (4, 64, 57333 * 64, 17080 * 64, 1000):
        for world in (1, 2, 3, 4, 8):
            spans = [owned_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a0 <= a1
            assert all(lo % 4 == 0 for lo, _ in spans)
