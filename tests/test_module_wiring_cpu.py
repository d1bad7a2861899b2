"""CPU check of the module plumbing in glass_b200/models.py (constructor wiring, layer order, JK concat, pooling
dispatch, label handling) with every `glass_b200.ops` entry point it uses replaced by an oracle function.
No kernel runs here -- the kernels are covered by the `-m gpu` tests; this guards refactors of the Python side
against the goldens of the unmodified reference."""
import pytest
import torch
import torch.nn.functional as F

from oracle import glass_oracle as O
from tests.helpers import MODEL_CASES, build_product_model, load_model_case, rel_err


class _Adj:
    def __init__(self, ei, ew, n, aggr):
        self.m = O.build_adj(ei, ew, n, aggr)
        self.n = n

    def __matmul__(self, x):
        return self.m @ x


def _act(x, act):
    from glass_b200._lib import ACT_ELU, ACT_RELU
    return F.elu(x) if act == ACT_ELU else F.relu(x) if act == ACT_RELU else x


def _pair(a1, a2, w0, b0, w1, b1, mask, z, act, path=None):
    a = a1 if a2 is None else torch.cat((a1, a2), dim=-1)
    p0, p1 = _act(F.linear(a, w0, b0), act), _act(F.linear(a, w1, b1), act)
    m = mask.reshape(-1, 1).bool()
    return torch.where(m, z * p1 + (1 - z) * p0, z * p0 + (1 - z) * p1)


def _segment_pool(emb, pos, mode):
    batch, nodes = O.pad2batch(pos)
    return O.pool_nodes(emb[nodes], batch, mode, pos.shape[0])


@pytest.fixture()
def oracle_ops(monkeypatch):
    from glass_b200 import models, ops
    monkeypatch.setattr(ops, "build_csr", lambda ei, ew, n, aggr: _Adj(ei, ew, n, aggr))
    monkeypatch.setattr(ops, "spmm", lambda adj, x: adj @ x)
    monkeypatch.setattr(ops, "pair_linear_mix", _pair)
    monkeypatch.setattr(ops, "graph_norm", lambda x, w, b, a, eps=1e-5, act=0, p=0.0, training=False:
                        _act(O.graph_norm(x, w, b, a, eps), act))
    monkeypatch.setattr(ops, "spmm_graph_norm", lambda adj, x, w, b, a, eps=1e-5, act=0, p=0.0, training=False:
                        _act(O.graph_norm(adj @ x, w, b, a, eps), act))
    monkeypatch.setattr(ops, "graph_norm_cat", lambda xs, w, b, a, eps=1e-5: O.graph_norm(torch.cat(list(xs), -1), w, b, a, eps))
    monkeypatch.setattr(ops, "graph_norm_pool", lambda x, w, b, a, eps, pos, mode:
                        _segment_pool(O.graph_norm(x, w, b, a, eps), pos, mode))
    monkeypatch.setattr(ops, "graph_norm_pool_cat", lambda xs, w, b, a, eps, pos, mode:
                        _segment_pool(O.graph_norm(torch.cat(list(xs), -1), w, b, a, eps), pos, mode))
    monkeypatch.setattr(ops, "label_mask", lambda z: (z > 0.5).to(torch.uint8))
    monkeypatch.setattr(ops, "embedding", lambda ids, table: table[ids])
    monkeypatch.setattr(ops, "segment_pool", _segment_pool)
    monkeypatch.setattr(ops, "segment_pool_batch", lambda x, batch, mode, size=None: O.pool_nodes(x, batch, mode, size))
    monkeypatch.setattr(ops, "pad2batch", O.pad2batch)
    models._adj_cache.clear()
    yield
    models._adj_cache.clear()


@pytest.mark.parametrize("name", MODEL_CASES)
def test_module_plumbing_reproduces_reference_goldens(name, oracle_ops):
    c = load_model_case(name)
    m = build_product_model(c["raw"], c["x"].shape[0])
    m.load_state_dict(c["sd"])
    m.eval()
    emb = m.NodeEmb(c["x"], c["ei"], c["ew"], c["z"])
    assert rel_err(emb.detach(), c["emb"]) < 1e-5
    pooled = m.Pool(emb, c["pos"], m.pools[0])
    assert rel_err(pooled.detach(), c["pooled"]) < 1e-5
    logits = m(c["x"], c["ei"], c["ew"], c["pos"], c["z"])
    assert rel_err(logits.detach(), c["logits"]) < 1e-5
    loss = O.loss_fn_for(c["cfg"].out_dim == 1)(logits, c["y"])
    assert abs(float(loss) - c["loss"]) < 1e-5 * max(1.0, abs(c["loss"]))
    loss.backward()
    for k, p in m.named_parameters():
        assert rel_err(p.grad, c["grads"][k]) < 2e-2 if "emb" in k else rel_err(p.grad, c["grads"][k]) < 1e-4, k


def test_generic_pool_module_path_uses_pad2batch_and_gather(oracle_ops):
    """A pool with a trans_fn cannot take the fused padded path: GLASS.Pool falls back to pad2batch + gather +
    PoolModule.forward (impl/models.py:346-350)."""
    import torch.nn as nn

    from glass_b200 import models
    c = load_model_case("ppibp_like")
    m = build_product_model(c["raw"], c["x"].shape[0])
    m.load_state_dict(c["sd"])
    m.eval()
    emb = m.NodeEmb(c["x"], c["ei"], c["ew"], c["z"])
    fused = m.Pool(emb, c["pos"], models.AddPool())
    generic = m.Pool(emb, c["pos"], models.AddPool(trans_fn=nn.Identity()))
    assert torch.allclose(fused, generic, rtol=1e-6, atol=1e-6)
    size = m.Pool(emb, c["pos"], models.SizePool(trans_fn=nn.Identity()))
    assert torch.allclose(size, m.Pool(emb, c["pos"], models.SizePool()), rtol=1e-6, atol=1e-6)
