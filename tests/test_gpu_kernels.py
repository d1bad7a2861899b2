"""GPU: kernel-level parity of the CUDA path (through the C ABI) against the CPU oracle.

Integer / index / normalisation results must be bit-exact; floating-point results within
1e-4 norm-wise relative error (BASELINE.json north_star)."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import glass_oracle as O
from tests.helpers import GOLDEN, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-4


@pytest.fixture(scope="module", autouse=True)
def _lib():
    from glass_b200 import build
    build.build()
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def rand_graph(n, n_und, seed):
    from glass_b200 import datasets
    e = datasets.uniform_edges(n, n_und, seed)
    return datasets.coalesce_undirected(e, torch.ones(e.shape[1]), n)


# ------------------------------------------------------------------------------------------ csr_build
def _check_csr(adj, ref, exact_val=True):
    assert np.array_equal(adj.rowptr.cpu().numpy(), ref["rowptr"])
    assert np.array_equal(adj.col.cpu().numpy(), ref["col"])
    assert np.array_equal(adj.rowptr_t.cpu().numpy(), ref["rowptr_t"])
    assert np.array_equal(adj.col_t.cpu().numpy(), ref["col_t"])
    v, vt, d = adj.val.cpu().numpy(), adj.val_t.cpu().numpy(), adj.deg.cpu().numpy()
    assert np.array_equal(v.view(np.uint32), ref["val"].view(np.uint32)), np.abs(v - ref["val"]).max()
    assert np.array_equal(vt.view(np.uint32), ref["val_t"].view(np.uint32))
    assert np.array_equal(d.view(np.uint32), ref["deg"].view(np.uint32))


@pytest.mark.parametrize("case", ["sym_unit", "unsorted_dup_selfloop", "tiny_weights", "single"])
@pytest.mark.parametrize("aggr", ["mean", "sum", "gcn"])
def test_csr_build_golden_cases_bit_exact(case, aggr):
    from glass_b200 import ops
    d = np.load(os.path.join(GOLDEN, "buildadj.npz"))
    ei, ew, n = d[f"{case}.ei"], d[f"{case}.ew"], int(d[f"{case}.n"])
    adj = ops.build_csr(torch.from_numpy(ei).to(DEV), torch.from_numpy(ew).to(DEV), n, aggr)
    _check_csr(adj, O.build_csr_numpy(ei, ew, n, aggr))          # oracle: bit-exact incl. non-unit weights
    idx, val = d[f"{case}.{aggr}.idx"], d[f"{case}.{aggr}.val"]    # reference .coalesce()
    assert np.array_equal(adj.indices().cpu().numpy(), idx)
    if case in ("sym_unit", "single"):
        assert np.array_equal(adj.values().cpu().numpy().view(np.uint32), val.view(np.uint32))
    else:
        np.testing.assert_allclose(adj.values().cpu().numpy(), val, rtol=1e-6, atol=0)


def test_csr_build_shipped_graphs_match_reference_digests():
    from glass_b200 import datasets, ops
    with open(os.path.join(GOLDEN, "buildadj_shipped.json")) as f:
        dig = json.load(f)
    for name, ref in dig.items():
        ei, ew, n = datasets.load_edges(name)
        for aggr in ("mean", "sum", "gcn"):
            adj = ops.build_csr(ei.to(DEV), ew.to(DEV), n, aggr)
            assert sha(adj.rowptr.cpu().numpy()) == ref[aggr]["rowptr"], (name, aggr)
            assert sha(adj.col.cpu().numpy()) == ref[aggr]["col"], (name, aggr)
            assert sha(adj.val.cpu().numpy()) == ref[aggr]["val"], (name, aggr)


@pytest.mark.parametrize("aggr", ["mean", "gcn"])
def test_csr_build_large_random_and_unsorted(aggr):
    from glass_b200 import ops
    ei, ew = rand_graph(20000, 300000, 3)
    ref = O.build_csr_numpy(ei.numpy(), ew.numpy(), 20000, aggr)
    _check_csr(ops.build_csr(ei.to(DEV), ew.to(DEV), 20000, aggr), ref)
    perm = torch.randperm(ei.shape[1], generator=torch.Generator().manual_seed(1))
    _check_csr(ops.build_csr(ei[:, perm].contiguous().to(DEV), ew[perm].contiguous().to(DEV), 20000, aggr), ref)


def test_csr_build_rejects_bad_input():
    from glass_b200 import ops
    ei = torch.tensor([[0, 5], [1, 2]], device=DEV)
    with pytest.raises(RuntimeError, match="outside"):
        ops.build_csr(ei, torch.ones(2, device=DEV), 4, "sum")
    with pytest.raises(NotImplementedError):
        ops.build_csr(ei, torch.ones(2, device=DEV), 8, "max")


def test_csr_empty_graph():
    from glass_b200 import ops
    adj = ops.build_csr(torch.zeros(2, 0, dtype=torch.int64, device=DEV), torch.zeros(0, device=DEV), 5, "mean")
    assert adj.nnz == 0 and adj.rowptr.tolist() == [0] * 6 and adj.deg.tolist() == [1.0] * 5
    y = ops.spmm(adj, torch.randn(5, 8, device=DEV))
    assert torch.count_nonzero(y) == 0


# ------------------------------------------------------------------------------------------ spmm
@pytest.mark.parametrize("h", [1, 4, 8, 17, 20, 32, 64, 100, 128, 200, 256])
def test_spmm_matches_oracle(h):
    from glass_b200 import ops
    n = 3000
    ei, ew = rand_graph(n, 40000, h)
    ew = torch.rand(ei.shape[1], generator=torch.Generator().manual_seed(h)) + 0.5
    adj = ops.build_csr(ei.to(DEV), ew.to(DEV), n, "mean")
    x = torch.randn(n, h, generator=torch.Generator().manual_seed(1))
    ref = O.build_adj(ei, ew, n, "mean") @ x
    y = ops.spmm(adj, x.to(DEV)).cpu()
    assert rel_err(y, ref) < 1e-5
    # backward = transposed CSR
    gy = torch.randn(n, h, generator=torch.Generator().manual_seed(2))
    ref_gx = O.build_adj(ei, ew, n, "mean").t() @ gy
    xg = x.to(DEV).requires_grad_(True)
    ops.spmm(adj, xg).backward(gy.to(DEV))
    assert rel_err(xg.grad.cpu(), ref_gx) < 1e-5


@pytest.mark.parametrize("h", [4, 8, 20, 32, 64])
def test_spmm_both_lane_configurations_match_oracle(h):
    """n = 3000 (above) runs the neighbour-parallel configuration (32/G slots per row), n = 20000 the
    one-chain-per-output throughput configuration (spmm.cu latency_regime)."""
    from glass_b200 import ops
    n = 20000
    ei, ew = rand_graph(n, 150000, h)
    adj = ops.build_csr(ei.to(DEV), ew.to(DEV), n, "gcn")
    x = torch.randn(n, h, generator=torch.Generator().manual_seed(1))
    ref = O.build_adj(ei, ew, n, "gcn")
    assert rel_err(ops.spmm(adj, x.to(DEV)).cpu(), ref @ x) < 1e-5
    assert rel_err(ops.spmm(adj.t(), x.to(DEV)).cpu(), ref.t() @ x) < 1e-5


def test_spmm_is_deterministic_and_sequential_order():
    """Single fp32 FMA chain per output in CSR order: equals a sequential CPU loop bit for bit
    whenever the products are exact (small integers)."""
    from glass_b200 import ops
    n, h = 500, 64
    ei, _ = rand_graph(n, 6000, 9)
    ew = torch.ones(ei.shape[1])
    adj = ops.build_csr(ei.to(DEV), ew.to(DEV), n, "sum")
    x = torch.randint(-8, 9, (n, h), generator=torch.Generator().manual_seed(0)).float()
    y1 = ops.spmm(adj, x.to(DEV))
    y2 = ops.spmm(adj, x.to(DEV))
    assert torch.equal(y1, y2)
    assert torch.equal(y1.cpu(), O.build_adj(ei, ew, n, "sum") @ x)


def test_spmm_skewed_rows_and_isolated_nodes():
    from glass_b200 import datasets, ops
    n = 5000
    e = datasets.powerlaw_edges(n, 60000, 0)
    e = e[:, (e[0] != 7) & (e[1] != 7)]                       # node 7 isolated
    ei, ew = datasets.coalesce_undirected(e, torch.ones(e.shape[1]), n)
    for aggr in ("mean", "sum", "gcn"):
        adj = ops.build_csr(ei.to(DEV), ew.to(DEV), n, aggr)
        x = torch.randn(n, 64, generator=torch.Generator().manual_seed(3))
        y = ops.spmm(adj, x.to(DEV)).cpu()
        assert rel_err(y, O.build_adj(ei, ew, n, aggr) @ x) < 1e-5
        assert torch.count_nonzero(y[7]) == 0


@pytest.mark.parametrize("h", [8, 17, 64, 128])
def test_spmm_row_split_plan_matches_plain_kernel(h):
    """Rows longer than the split length go through the planned kernel (chunk partials + ordered combine)."""
    from glass_b200 import datasets, ops
    n = 4000
    e = datasets.powerlaw_edges(n, 50000, 1)
    ei, ew = datasets.coalesce_undirected(e, torch.ones(e.shape[1]), n)
    adj = ops.build_csr(ei.to(DEV), ew.to(DEV), n, "gcn")
    x = torch.randn(n, h, generator=torch.Generator().manual_seed(2)).to(DEV)
    gy = torch.randn(n, h, generator=torch.Generator().manual_seed(3)).to(DEV)
    adj.plan = adj.plan_t = None
    xg = x.clone().requires_grad_(True)
    y0 = ops.spmm(adj, xg)
    y0.backward(gy)
    adj.make_plans(max_len=32)
    assert adj.plan is not None and adj.plan.n_long > 10 and adj.plan.n_items > n
    xs = x.clone().requires_grad_(True)
    y1 = ops.spmm(adj, xs)
    y1.backward(gy)
    assert rel_err(y1.detach().cpu(), y0.detach().cpu()) < 5e-6   # chunked summation order
    assert rel_err(xs.grad.cpu(), xg.grad.cpu()) < 5e-6
    ref = O.build_adj(ei, ew, n, "gcn") @ x.cpu()
    assert rel_err(y1.detach().cpu(), ref) < 1e-5
    # uniform graphs need no plan
    ei2, ew2 = rand_graph(2000, 20000, 5)
    assert ops.build_csr(ei2.to(DEV), ew2.to(DEV), 2000, "sum").plan is None


@pytest.mark.parametrize("h,planned", [(64, False), (64, True), (17, False), (8, True)])
def test_spmm_accumulate_column_phases(h, planned):
    """y (+)= A x (glass_spmm_csr_acc): the product taken as column-partitioned phases -- first the entries whose
    column is < n/3, then the middle third, then the rest, each accumulating into y -- equals adj @ x (oracle,
    impl/models.py:164).  This is how the row-partitioned SpMM consumes the peers' feature shards as they arrive."""
    from glass_b200 import datasets, ops
    from glass_b200.partition import split_columns_by_owner
    n = 3000
    e = datasets.powerlaw_edges(n, 40000, 7)
    ei, ew = datasets.coalesce_undirected(e, torch.ones(e.shape[1]), n)
    adj = ops.build_csr(ei.to(DEV), ew.to(DEV), n, "gcn")
    x = torch.randn(n, h, generator=torch.Generator().manual_seed(8)).to(DEV)
    bounds = [0, 1000, 2000, 3000]
    subs = split_columns_by_owner(adj.rowptr, adj.col, adj.val, bounds, rebase=False)
    y = torch.full((n, h), 7.0, device=DEV)                      # the first phase overwrites
    for i, (rp, c, v) in enumerate(subs):
        plan = None
        if planned:
            p = ops.RowSplitPlan(rp, 32)
            plan = p if p.n_long else None
        if i == 0:
            ops._run_spmm(rp, c, v, plan, x, y)
        else:
            ops._run_spmm(rp, c, v, plan, x, y, accumulate=True)
    assert rel_err(y.cpu(), O.build_adj(ei, ew, n, "gcn") @ x.cpu()) < 1e-5


# ------------------------------------------------------------------------------------------ pair GEMM
def _pair_ref(a, w0, b0, w1, b1, mask, z, act):
    f = {0: (lambda t: t), 1: torch.relu, 2: torch.nn.functional.elu}[act]
    p0 = f(torch.nn.functional.linear(a, w0, b0))
    p1 = f(torch.nn.functional.linear(a, w1, b1))
    return torch.where(mask.bool().view(-1, 1), z * p1 + (1 - z) * p0, z * p0 + (1 - z) * p1)   # models.py:161


@pytest.mark.parametrize("path", ["simt", "auto"])
@pytest.mark.parametrize("n,k1,k2,h,act,z", [(1000, 64, 0, 64, 2, 0.8), (777, 64, 64, 64, 0, 0.75),
                                            (513, 8, 8, 8, 0, 1.0), (300, 17, 0, 17, 2, 0.9),
                                            (300, 17, 17, 17, 0, 0.9), (257, 20, 20, 20, 1, 0.95),
                                            (4096, 128, 0, 128, 2, 0.6), (130, 64, 64, 32, 2, 0.8),
                                            (5, 16, 0, 16, 1, 0.5)])
def test_pair_linear_mix_fwd_bwd(n, k1, k2, h, act, z, path):
    from glass_b200 import _lib, ops
    g = torch.Generator().manual_seed(n + h)
    a1 = torch.randn(n, k1, generator=g)
    a2 = torch.randn(n, k2, generator=g) if k2 else None
    k = k1 + k2
    w0, w1 = torch.randn(h, k, generator=g) / k ** 0.5, torch.randn(h, k, generator=g) / k ** 0.5
    b0, b1 = torch.randn(h, generator=g), torch.randn(h, generator=g)
    mask = (torch.rand(n, generator=g) > 0.5).to(torch.uint8)
    gout = torch.randn(n, h, generator=g)
    # oracle in fp64 (the reference computes fp32; both are compared against the exact value)
    cpu = [t.double().requires_grad_(True) if t is not None else None for t in (a1, a2, w0, b0, w1, b1)]
    a_cat = cpu[0] if a2 is None else torch.cat((cpu[0], cpu[1]), dim=-1)
    ref = _pair_ref(a_cat, cpu[2], cpu[3], cpu[4], cpu[5], mask, z, act)
    ref.backward(gout.double())
    dev = [t.to(DEV).requires_grad_(True) if t is not None else None for t in (a1, a2, w0, b0, w1, b1)]
    pid = {"simt": _lib.GEMM_SIMT, "auto": _lib.GEMM_AUTO}[path]
    out = ops.pair_linear_mix(dev[0], dev[1], dev[2], dev[3], dev[4], dev[5], mask.to(DEV), z, act, pid)
    assert rel_err(out.cpu(), ref) < 2e-5
    out.backward(gout.to(DEV))
    for name, c, d in zip(("a1", "a2", "w0", "b0", "w1", "b1"), cpu, dev):
        if c is not None:
            assert rel_err(d.grad.cpu(), c.grad) < 5e-5, name


# ------------------------------------------------------------------------------------------ GraphNorm
@pytest.mark.parametrize("n,c,act,p", [(1000, 64, 0, 0.0), (999, 64, 2, 0.5), (300, 17, 0, 0.3), (257, 20, 1, 0.05),
                                       (5000, 128, 2, 0.0), (64, 8, 0, 0.0), (2000, 256, 2, 0.2),
                                       (20000, 17, 2, 0.3), (30000, 8, 1, 0.1), (12000, 20, 0, 0.0),
                                       (600, 256, 2, 0.2), (5000, 8, 2, 0.0), (7, 4, 0, 0.0),
                                       (180, 256, 2, 0.2), (700, 64, 2, 0.5)])
def test_graph_norm_fwd_bwd(n, c, act, p):
    """Matrices up to 48 K elements take the one-launch cluster kernel, larger ones the three-kernel path
    (glass_graphnorm_launches); both are covered for vector (c % 4 == 0) and scalar column layouts."""
    from glass_b200 import ops
    g = torch.Generator().manual_seed(n + c)
    x = torch.randn(n, c, generator=g) * 3 + torch.randn(1, c, generator=g) * 5   # non-zero column means
    w, b, a = (torch.randn(c, generator=g) for _ in range(3))
    a = a * 0.3 + 1.0
    keep = (torch.rand(n, c, generator=g) >= p).to(torch.uint8) if p > 0 else None
    gout = torch.randn(n, c, generator=g)
    cpu = [t.clone().requires_grad_(True) for t in (x, w, b, a)]
    f = {0: (lambda t: t), 1: torch.relu, 2: torch.nn.functional.elu}[act]
    ref = f(O.graph_norm(*cpu))
    if keep is not None:
        ref = ref * keep.float() / (1 - p)
    ref.backward(gout)
    dev = [t.to(DEV).requires_grad_(True) for t in (x, w, b, a)]
    if keep is not None:
        with ops.inject_keep_masks([keep.to(DEV)]):
            out = ops.graph_norm(*dev, 1e-5, act, p, True)
    else:
        out = ops.graph_norm(*dev, 1e-5, act, p, True)
    assert rel_err(out.cpu(), ref) < 2e-5
    out.backward(gout.to(DEV))
    for name, cc, d in zip(("x", "weight", "bias", "mean_scale"), cpu, dev):
        assert rel_err(d.grad.cpu(), cc.grad) < TOL, name


def test_graph_norm_dropout_in_kernel_generator():
    """Philox bits generated inside the kernels: right drop rate, survivors scaled by 1/(1-p), a fresh mask
    per call, and backward regenerates exactly the mask forward used (checked against the injected-mask path)."""
    from glass_b200 import ops
    g = torch.Generator().manual_seed(0)
    for n, c, act, p in ((4000, 64, 0, 0.5), (777, 17, 2, 0.3)):
        x = (torch.randn(n, c, generator=g) + 0.1).to(DEV)
        w, b, a = (torch.rand(c, generator=g) + 0.5).to(DEV), torch.randn(c, generator=g).to(DEV), torch.ones(c, device=DEV)
        ev = ops.graph_norm(x, w, b, a, 1e-5, act, p, False)
        xs = [x.clone().requires_grad_(True) for _ in range(2)]
        ws = [w.clone().requires_grad_(True) for _ in range(2)]
        out = ops.graph_norm(xs[0], ws[0], b, a, 1e-5, act, p, True)
        out2 = ops.graph_norm(x, w, b, a, 1e-5, act, p, True)
        keep = (out != 0) | (ev == 0)
        frac = float((~keep).float().mean())
        assert abs(frac - p) < 0.02, frac
        assert rel_err(out[keep].detach(), ev[keep] / (1 - p)) < 1e-6
        assert not torch.equal(out2 != 0, out != 0)                      # next call, next mask
        gout = torch.randn(n, c, generator=g).to(DEV)
        out.backward(gout)
        with ops.inject_keep_masks([keep.to(torch.uint8)]):
            ref = ops.graph_norm(xs[1], ws[1], b, a, 1e-5, act, p, True)
        assert torch.equal(ref, out)
        ref.backward(gout)
        assert rel_err(xs[0].grad, xs[1].grad) < 1e-6 and rel_err(ws[0].grad, ws[1].grad) < 1e-6
    # re-seeding repeats the sequence
    ops.manual_seed(123)
    m1 = ops.graph_norm(x, w, b, a, 1e-5, 0, 0.5, True) != 0
    ops.manual_seed(123)
    m2 = ops.graph_norm(x, w, b, a, 1e-5, 0, 0.5, True) != 0
    assert torch.equal(m1, m2)


def test_graph_norm_cat_matches_norm_of_concat():
    from glass_b200 import ops
    g = torch.Generator().manual_seed(5)
    xs = [torch.randn(700, 64, generator=g) + i for i in range(3)]
    w, b, a = (torch.randn(192, generator=g) for _ in range(3))
    gout = torch.randn(700, 192, generator=g)
    cpu_xs = [t.clone().requires_grad_(True) for t in xs]
    cw, cb, ca = (t.clone().requires_grad_(True) for t in (w, b, a))
    ref = O.graph_norm(torch.cat(cpu_xs, dim=-1), cw, cb, ca)
    ref.backward(gout)
    dxs = [t.to(DEV).requires_grad_(True) for t in xs]
    dw, db, da = (t.to(DEV).requires_grad_(True) for t in (w, b, a))
    out = ops.graph_norm_cat(dxs, dw, db, da)
    assert rel_err(out.cpu(), ref) < 2e-5
    out.backward(gout.to(DEV))
    for c, d in zip(cpu_xs + [cw, cb, ca], dxs + [dw, db, da]):
        assert rel_err(d.grad.cpu(), c.grad) < TOL


# ------------------------------------------------------------------------------------------ embedding / pooling / labels
@pytest.mark.parametrize("kind", ["one", "arange", "random"])
@pytest.mark.parametrize("h", [8, 17, 64])
def test_embedding_fwd_bwd(kind, h):
    from glass_b200 import ops
    n, rows = 3001, {"one": 2, "arange": 3001, "random": 50}[kind]
    g = torch.Generator().manual_seed(h)
    ids = {"one": torch.ones(n, dtype=torch.int64), "arange": torch.arange(n),
           "random": torch.randint(0, 50, (n,), generator=g)}[kind]
    table = torch.randn(rows, h, generator=g)
    gout = torch.randn(n, h, generator=g)
    ct = table.clone().requires_grad_(True)
    ref = torch.nn.functional.embedding(ids, ct)
    ref.backward(gout)
    dt = table.to(DEV).requires_grad_(True)
    out = ops.embedding(ids.to(DEV), dt)
    assert torch.equal(out.cpu(), ref.detach())
    out.backward(gout.to(DEV))
    assert rel_err(dt.grad.cpu(), ct.grad) < 1e-5


def _rand_pad(n, b, lmax, seed, empty_row=None):
    g = np.random.default_rng(seed)
    pad = -np.ones((b, lmax), dtype=np.int64)
    for i in range(b):
        k = int(g.integers(1, lmax + 1))
        pad[i, :k] = g.choice(n, size=k, replace=False)
    if empty_row is not None:
        pad[empty_row, :] = -1
    return torch.from_numpy(pad)


@pytest.mark.parametrize("mode", ["sum", "mean", "max", "size"])
@pytest.mark.parametrize("d", [8, 17, 128, 320])
def test_segment_pool_fwd_bwd(mode, d):
    from glass_b200 import ops
    n, b, lmax = 900, 13, 40
    pad = _rand_pad(n, b, lmax, d)
    g = torch.Generator().manual_seed(d)
    emb = torch.randn(n, d, generator=g)
    gout = torch.randn(b, d, generator=g)
    ce = emb.clone().requires_grad_(True)
    batch, pos = O.pad2batch(pad)
    ref = O.pool_nodes(ce[pos], batch, mode, n_seg=b)
    ref.backward(gout)
    de = emb.to(DEV).requires_grad_(True)
    out = ops.segment_pool(de, pad.to(DEV), mode)
    assert rel_err(out.cpu(), ref) < 1e-5
    out.backward(gout.to(DEV))
    assert rel_err(de.grad.cpu(), ce.grad) < 1e-5
    # PoolModule.forward(x, batch) variant on gathered rows
    dx = emb[pos].to(DEV).requires_grad_(True)
    out2 = ops.segment_pool_batch(dx, batch.to(DEV), mode)
    assert rel_err(out2.cpu(), ref) < 1e-5
    out2.backward(gout.to(DEV))
    cx = emb[pos].clone().requires_grad_(True)
    O.pool_nodes(cx, batch, mode, n_seg=b).backward(gout)
    assert rel_err(dx.grad.cpu(), cx.grad) < 1e-5


def test_segment_pool_empty_middle_row_gives_zero():
    from glass_b200 import ops
    pad = _rand_pad(100, 5, 6, 0, empty_row=2)
    emb = torch.randn(100, 16)
    for mode in ("sum", "mean", "max", "size"):
        out = ops.segment_pool(emb.to(DEV), pad.to(DEV), mode).cpu()
        assert torch.count_nonzero(out[2]) == 0 and torch.isfinite(out).all()


def test_labels_and_pad2batch_known_answers():
    from glass_b200 import utils
    d = np.load(os.path.join(GOLDEN, "utils_kat.npz"))
    for tag, n in (("doc", 9), ("rand", 500)):
        pad = torch.from_numpy(d[f"{tag}_pad"]).to(DEV)
        b, p = utils.pad2batch(pad)
        assert np.array_equal(b.cpu().numpy(), d[f"{tag}_batch"])
        assert np.array_equal(p.cpu().numpy(), d[f"{tag}_pos"])
        z = utils.MaxZOZ(torch.zeros(n, 1, device=DEV), pad)
        assert z.dtype == torch.int64 and np.array_equal(z.cpu().numpy(), d[f"{tag}_z"])
    from glass_b200 import ops
    z, m = ops.maxzoz(500, torch.from_numpy(d["rand_pad"]).to(DEV), with_mask=True)
    assert torch.equal(m.cpu(), (z.cpu() > 0.5).to(torch.uint8))
    assert torch.equal(ops.label_mask(z).cpu(), m.cpu())
    big = _rand_pad(60000, 300, 400, 1)
    b, p = utils.pad2batch(big.to(DEV))
    rb, rp = O.pad2batch(big)
    assert torch.equal(b.cpu(), rb) and torch.equal(p.cpu(), rp)


@pytest.mark.parametrize("mode", ["sum", "mean", "size"])
@pytest.mark.parametrize("n,d,b,lmax", [(3000, 64, 6, 40), (777, 17, 5, 9), (5000, 8, 3, 20), (40000, 128, 40, 70)])
def test_graph_norm_pool_matches_norm_then_pool(mode, n, d, b, lmax):
    """ops.graph_norm_pool == pool(GraphNorm(x)[subG_node]) of the oracle (impl/models.py:266 -> :346-350), forward and
    every gradient (x, weight, bias, mean_scale).  Nodes shared by several subgraphs, a node listed twice in one
    subgraph and an empty subgraph are part of the batch."""
    from glass_b200 import ops
    g = torch.Generator().manual_seed(n + d)
    x = torch.randn(n, d, generator=g) * 2 + torch.randn(1, d, generator=g) * 3
    w, bias, ms = torch.randn(d, generator=g), torch.randn(d, generator=g), torch.rand(d, generator=g) + 0.5
    pos = _rand_pad(n, b, lmax, seed=d, empty_row=b - 1)
    pos[1, :3] = pos[0, :3]                       # nodes shared by two subgraphs
    pos[2, 1] = pos[2, 0]                         # a node listed twice in one subgraph
    gout = torch.randn(b, d, generator=g)
    cpu = [t.clone().requires_grad_(True) for t in (x, w, bias, ms)]
    batch, nodes = O.pad2batch(pos)
    ref = O.pool_nodes(O.graph_norm(*cpu)[nodes], batch, mode, b)
    ref.backward(gout)
    dev = [t.to(DEV).requires_grad_(True) for t in (x, w, bias, ms)]
    out = ops.graph_norm_pool(*dev, 1e-5, pos.to(DEV), mode)
    assert rel_err(out.detach().cpu(), ref.detach()) < 2e-5
    out.backward(gout.to(DEV))
    for name, c_, d_ in zip(("x", "weight", "bias", "mean_scale"), cpu, dev):
        assert rel_err(d_.grad.cpu(), c_.grad) < TOL, name
    # and bit-for-bit the same pooled values as the two separate operators of the product
    two = ops.segment_pool(ops.graph_norm(*[t.detach() for t in dev], 1e-5, 0, 0.0, False), pos.to(DEV), mode)
    assert rel_err(out.detach(), two) < 1e-6


@pytest.mark.parametrize("mode", ["sum", "mean", "size"])
@pytest.mark.parametrize("widths", [(64, 64), (20, 20), (8, 32, 16)])
def test_graph_norm_pool_cat_matches_norm_of_concat_then_pool(mode, widths):
    """ops.graph_norm_pool_cat == pool(GraphNorm(cat(xs))[subG_node]) (the JK case, impl/models.py:263-267 -> :346-350):
    forward and the gradients of every block and of the norm's parameters, against the oracle."""
    from glass_b200 import ops
    n, b, lmax = 2500, 7, 30
    g = torch.Generator().manual_seed(sum(widths))
    d = sum(widths)
    xs = [torch.randn(n, w_, generator=g) * 2 + torch.randn(1, w_, generator=g) for w_ in widths]
    w, bias, ms = torch.randn(d, generator=g), torch.randn(d, generator=g), torch.rand(d, generator=g) + 0.5
    pos = _rand_pad(n, b, lmax, seed=d, empty_row=b - 1)
    pos[1, :3] = pos[0, :3]
    gout = torch.randn(b, d, generator=g)
    cpu = [t.clone().requires_grad_(True) for t in (*xs, w, bias, ms)]
    batch, nodes = O.pad2batch(pos)
    ref = O.pool_nodes(O.graph_norm(torch.cat(cpu[:len(xs)], -1), *cpu[len(xs):])[nodes], batch, mode, b)
    ref.backward(gout)
    dev = [t.to(DEV).requires_grad_(True) for t in (*xs, w, bias, ms)]
    out = ops.graph_norm_pool_cat(dev[:len(xs)], *dev[len(xs):], 1e-5, pos.to(DEV), mode)
    assert rel_err(out.detach().cpu(), ref.detach()) < 2e-5
    out.backward(gout.to(DEV))
    for i, (c_, d_) in enumerate(zip(cpu, dev)):
        assert rel_err(d_.grad.cpu(), c_.grad) < TOL, i


# ------------------------------------------------------------------------------------------ row partitioning
@pytest.mark.parametrize("overlap", [False, True, "pipelined"])
@pytest.mark.parametrize("world", [1, 3, 4, 8])
def test_row_partitioned_spmm_single_device_emulation(world, overlap):
    """All ranks' blocks built on one GPU; the all-gather is emulated by concatenating the padded shards.
    overlap=True multiplies the locally-owned columns first (while the gather would be in flight); "pipelined" takes
    the product in one accumulating phase per source rank (own shard first, then rank+1, ...)."""
    from glass_b200 import datasets, ops
    from glass_b200.partition import RowPartitionedAdj
    n, h = 6000, 64
    e = datasets.powerlaw_edges(n, 80000, 2)
    ei, ew = datasets.coalesce_undirected(e, torch.ones(e.shape[1]), n)
    adj = ops.build_csr(ei.to(DEV), ew.to(DEV), n, "mean")
    x = torch.randn(n, h, generator=torch.Generator().manual_seed(1)).to(DEV)
    gy = torch.randn(n, h, generator=torch.Generator().manual_seed(2)).to(DEV)
    # reference op on the CPU (sparse COO @ dense, impl/models.py:164) -- not the CUDA path against itself
    xf = x.cpu().requires_grad_(True)
    y_ref = O.build_adj(ei, ew, n, "mean") @ xf
    y_ref.backward(gy.cpu())
    pipelined = overlap == "pipelined"
    parts = [RowPartitionedAdj(adj, r, world, overlap=overlap is True, pipelined=pipelined) for r in range(world)]
    nnz = [p.nnz_local for p in parts]
    if overlap is True:
        assert all(p.own.nnz + p.rem.nnz == p.nnz_local for p in parts)
    if pipelined:
        assert all(sum(ph.nnz for ph in p.phases) == p.nnz_local for p in parts)
    assert sum(nnz) == adj.nnz and max(nnz) <= 1.3 * adj.nnz / world + 4096      # balanced by entries
    pad = parts[0].pad

    def padded(t, p):
        out = torch.zeros(pad, t.shape[1], device=DEV)
        out[:p.rows] = t[p.lo:p.hi]
        return out

    x_full = torch.cat([padded(x, p) for p in parts])
    gy_full = torch.cat([padded(gy, p) for p in parts])
    for p in parts:
        xs = p.shard(x).requires_grad_(True)
        state = {"fwd": True}

        def fake_gather(send, state=state):
            full = x_full if state["fwd"] else gy_full
            state["fwd"] = False
            return full

        p.gather_override = fake_gather
        p.exchange_override = lambda send, fg=fake_gather: list(fg(send).split(pad))
        y = p.spmm(xs)
        assert rel_err(y.detach().cpu(), y_ref[p.lo:p.hi].detach()) < 1e-5
        y.backward(gy[p.lo:p.hi].contiguous())
        assert rel_err(xs.grad.cpu(), xf.grad[p.lo:p.hi]) < 1e-5


# ------------------------------------------------------------------------------------------ optimizer
def test_fused_adam_matches_torch_adam():
    from glass_b200.optim import FusedAdam
    g = torch.Generator().manual_seed(0)
    shapes = [(57, 64), (64,), (3, 5), (1,), (4099,), (128, 128)]
    p_ref = [torch.nn.Parameter(torch.randn(*s, generator=g).to(DEV)) for s in shapes]
    p_new = [torch.nn.Parameter(p.detach().clone()) for p in p_ref]
    ref = torch.optim.Adam(p_ref, lr=1e-2)
    new = FusedAdam(p_new, lr=1e-2)
    for step in range(6):
        for a, b in zip(p_ref, p_new):
            grad = torch.randn(a.shape, generator=g).to(DEV)
            a.grad, b.grad = grad.clone(), grad.clone()
        ref.step()
        new.step()
        if step == 3:
            for grp in ref.param_groups:
                grp["lr"] = 3e-3
            new.set_lr(3e-3)
    for s_, a, b in zip(shapes, p_ref, p_new):
        assert rel_err(b.detach().cpu(), a.detach().cpu()) < 2e-6, s_
    # a parameter without a gradient is left untouched (note: the step count is global, torch's is per parameter)
    before = [p.detach().clone() for p in p_new]
    for b in p_new:
        b.grad = torch.ones_like(b)
    p_new[2].grad = None
    new.step()
    assert torch.equal(p_new[2].detach(), before[2]) and not torch.equal(p_new[0].detach(), before[0])


# ------------------------------------------------------------------------------------------ to_undirected (device)
def test_to_undirected_on_device_matches_host_coalesce():
    """ops.to_undirected == PyG to_undirected + coalesce (datasets.coalesce_undirected restates it on the host and is
    pinned against the reference's loader in tests/test_host_logic.py): directed, duplicated, self-loop input;
    an already undirected list comes back untouched."""
    from glass_b200 import datasets, ops
    g = torch.Generator().manual_seed(3)
    n = 500
    e = torch.randint(0, n, (2, 4000), generator=g)
    e[:, :50] = e[:, 50:100]                                    # duplicates
    e[1, 100:120] = e[0, 100:120]                               # self loops
    w = torch.rand(e.shape[1], generator=g) + 0.5
    ref_i, ref_w = datasets.coalesce_undirected(e, w, n)
    got_i, got_w = ops.to_undirected(e.to(DEV), w.to(DEV), n)
    assert torch.equal(got_i.cpu(), ref_i)
    assert rel_err(got_w.cpu(), ref_w) < 1e-6
    again_i, again_w = ops.to_undirected(got_i, got_w, n)       # idempotent on an undirected, coalesced list
    assert again_i.data_ptr() == got_i.data_ptr() and torch.equal(again_w, got_w)
    dup = torch.tensor([[0, 0], [1, 1]], device=DEV)            # a duplicated directed edge is NOT "already undirected"
    di, dw = ops.to_undirected(dup, torch.ones(2, device=DEV), 4)
    assert di.tolist() == [[0, 1], [1, 0]] and dw.tolist() == [2.0, 2.0]
    # the graph container uses it when its tensors live on the GPU
    bg = datasets.BaseGraph(torch.empty((n, 1, 0), device=DEV), e.to(DEV), w.to(DEV), torch.zeros((1, 1), dtype=torch.int64),
                            torch.zeros(1), torch.zeros(1, dtype=torch.int64))
    assert torch.equal(bg.edge_index.cpu(), ref_i)
