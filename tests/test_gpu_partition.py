"""GPU: the row-partitioned model (glass_b200.partition.PartitionedGLASS) on ONE device.

All ranks run one after the other; every collective is emulated by replaying the contributions the other ranks
recorded in the previous sweep (a collective's inputs only depend on earlier collectives, so after as many sweeps
as there are collectives in sequence every rank sees exactly what a real all-gather / all-reduce would deliver).
The result must equal the replicated single-device model AND the oracle's goldens of the unmodified reference."""
import pytest
import torch

from oracle import glass_oracle as O
from tests.helpers import build_product_model, load_model_case, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module", autouse=True)
def _lib():
    from glass_b200 import build
    build.build()
    torch.cuda.set_device(0)


class ReplayComm:
    def __init__(self, world, rank, prev):
        self.world, self.rank, self.prev = world, rank, prev      # prev: {call index: [tensor per rank]}
        self.calls, self.record = 0, {}

    def _next(self, t):
        i = self.calls
        self.calls += 1
        self.record[i] = t.detach().clone()
        return self.prev.get(i)

    def all_reduce_sum(self, t):
        others = self._next(t)
        if others is not None:
            tot = torch.zeros_like(t)
            for r in range(self.world):
                tot += t if r == self.rank else others[r]
            t.copy_(tot)
        return t

    def all_gather(self, send):
        others = self._next(send)
        parts = [send if r == self.rank else (others[r] if others is not None else torch.zeros_like(send))
                 for r in range(self.world)]
        return torch.cat(parts, dim=0)


def _run_partitioned(case, world, p_drop=0.0, keeps=None, pipelined=False):
    """Sweeps until the recorded collectives stop changing; returns (logits of rank 0, summed parameter gradients)."""
    from glass_b200 import ops, utils
    from glass_b200.partition import PartitionedGLASS, RowPartitionedAdj
    c = case
    n = c["x"].shape[0]
    ei, ew, pos = c["ei"].to(DEV), c["ew"].to(DEV), c["pos"].to(DEV)
    y = c["y"].to(DEV)
    z = c["z"].to(DEV) if c["z"] is not None else None
    model = build_product_model(c["raw"], n, dropout=p_drop)
    model.load_state_dict(c["sd"])
    model = model.to(DEV).train()
    adj = ops.build_csr(ei, ew, n, c["raw"]["aggr"])
    parts = [RowPartitionedAdj(adj, r, world, pipelined=pipelined) for r in range(world)]
    table = model.conv.input_emb.weight
    ids = c["x"].reshape(-1).to(DEV)
    loss_fn = O.loss_fn_for(c["cfg"].out_dim == 1)
    prev, result = {}, None
    for sweep in range(60):
        records, outs = [], []
        for r in range(world):
            comm = ReplayComm(world, r, prev)
            parts[r].gather_override = comm.all_gather
            parts[r].exchange_override = lambda send, c_=comm, p_=parts[r]: list(c_.all_gather(send).split(p_.pad))
            pm = PartitionedGLASS(model, parts[r], comm)
            model.zero_grad(set_to_none=True)
            h_local = table[ids[parts[r].lo:parts[r].hi]]                    # this rank's rows of the input embedding
            ctx = ops.inject_keep_masks([k[parts[r].lo:parts[r].hi].contiguous() for k in keeps]) if keeps else None
            if ctx:
                with ctx:
                    logits = pm(h_local, pos, z)
            else:
                logits = pm(h_local, pos, z)
            loss_fn(logits, y).backward()
            grads = {k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in model.named_parameters()}
            head = {k for k, _ in model.named_parameters() if k.startswith("preds.")}
            outs.append((logits.detach().clone(), grads, head))
            records.append(comm.record)
        new_prev = {i: [records[r][i] for r in range(world)] for i in records[0]}
        same = bool(prev) and all(all(torch.allclose(a, b, rtol=1e-6, atol=1e-9) for a, b in zip(new_prev[i], prev[i]))
                                  for i in new_prev)
        prev = new_prev
        if same:
            result = outs
            break
    assert result is not None, "collective replay did not converge"
    total = {}
    for k in result[0][1]:
        gs = [o[1][k] for o in result if o[1][k] is not None]
        if not gs:
            continue
        scale = (1.0 / world) if k in result[0][2] else 1.0                    # reduce_grads(): head computed redundantly
        total[k] = sum(gs) * scale
    return result[0][0], total, model


@pytest.mark.parametrize("name,world,pipelined", [("ppibp_like", 3, False), ("emuser_like", 2, False),
                                                  ("cutratio_like", 4, False), ("ppibp_like", 3, True),
                                                  ("emuser_like", 4, True)])
def test_partitioned_model_matches_reference_golden(name, world, pipelined):
    c = load_model_case(name)
    logits, grads, model = _run_partitioned(c, world, pipelined=pipelined)
    assert rel_err(logits.cpu(), c["logits"]) < 1e-4
    table_key = "conv.input_emb.weight"
    for k, g in grads.items():
        if c["raw"]["emb"] == "one" and ("input_emb" in k or "emb_gn" in k):
            continue                                                          # degenerate zero-variance input
        assert rel_err(g.cpu(), c["grads"][k]) < 1e-4, k
    assert table_key in grads


def test_dist_graph_norm_single_rank_equals_graph_norm():
    from glass_b200 import ops
    from glass_b200.partition import Comm, _DistGraphNorm
    g = torch.Generator().manual_seed(0)
    n, c = 3000, 64
    x = (torch.randn(n, c, generator=g) * 2 + 1).to(DEV)
    w, b, a = (torch.randn(c, generator=g).to(DEV) for _ in range(3))
    keep = (torch.rand(n, c, generator=g) > 0.4).to(torch.uint8).to(DEV)
    gout = torch.randn(n, c, generator=g).to(DEV)
    outs = []
    for fn in ("ref", "dist"):
        xs = x.clone().requires_grad_(True)
        ps = [t.clone().requires_grad_(True) for t in (w, b, a)]
        with ops.inject_keep_masks([keep]):
            if fn == "ref":
                out = ops.graph_norm(xs, *ps, 1e-5, 2, 0.4, True)
            else:
                out = _DistGraphNorm.apply(xs, *ps, 1e-5, 2, 0.4, True, n, Comm())
        out.backward(gout)
        outs.append((out.detach(), xs.grad, *[p.grad for p in ps]))
    for a_, b_ in zip(*outs):
        assert rel_err(b_.cpu(), a_.cpu()) < 1e-6
