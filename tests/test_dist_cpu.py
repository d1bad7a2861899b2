"""CPU, world_size 2, gloo: host-side logic of label-batch data parallelism (no kernels involved)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from glass_b200.dist import shard_batches, shared_permutation
    from glass_b200.graphed import GradAverager
    torch.manual_seed(0)
    # one "big" tensor (reduced in place, like the N x H embedding table) and several small ones (flat buffer)
    model = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    GradAverager.BIG = 10
    params = list(model.parameters())
    avg = GradAverager(params)
    assert len(avg.big) == 1 and len(avg.small) == 3
    perm = shared_permutation(20, seed=7, epoch=1)
    mine = shard_batches(10, rank, world)
    data = torch.arange(80, dtype=torch.float32).reshape(20, 4)
    for b in mine[:2]:
        idx = perm[b * 2:(b + 1) * 2]
        for p in params:
            p.grad = None
        loss = model(data[idx]).sum()
        loss.backward()
        local = torch.cat([p.grad.flatten() for p in params])
        avg()                                   # the product path's averaging (graphed.GraphedTrainStep calls this)
        gathered = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        now = torch.cat([p.grad.flatten() for p in params])
        assert torch.allclose(now, sum(gathered) / world)
    out[rank] = (perm.tolist(), mine, now.tolist())
    dist.barrier()
    dist.destroy_process_group()


def test_label_batch_dp_gloo_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, 29531, out), nprocs=world, join=True)
    (p0, b0, g0), (p1, b1, g1) = out[0], out[1]
    assert p0 == p1                                     # one shared order, no communication
    assert b0 == [0, 2, 4, 6, 8] and b1 == [1, 3, 5, 7, 9]
    assert g0 == g1                                     # averaged gradients identical on both ranks


def _eval_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from glass_b200 import train
    from glass_b200.dist import sharded_test
    torch.manual_seed(3)
    model = torch.nn.Linear(4, 3)
    g = torch.Generator().manual_seed(1)
    sizes = [3, 1, 4, 2, 5]                                # odd batch count, ragged last batches
    batches = [(torch.randn(n, 4, generator=g), torch.randint(0, 3, (n,), generator=g)) for n in sizes]
    metric = lambda pred, y: float((pred.argmax(-1) == y).mean())
    loss_fn = torch.nn.CrossEntropyLoss()
    ref_score, ref_loss = train.test(model, batches, metric, loss_fn)
    score, loss = sharded_test(model, batches, metric, loss_fn)
    out[rank] = (ref_score, float(ref_loss), score, float(loss))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_eval_matches_single_device_eval_gloo_world2():
    world = 2
    out = mp.Manager().dict()
    mp.spawn(_eval_worker, args=(world, 29533, out), nprocs=world, join=True)
    for r in range(world):
        ref_score, ref_loss, score, loss = out[r]
        assert score == ref_score and abs(loss - ref_loss) < 1e-6


def test_shard_batches_drops_remainder():
    from glass_b200.dist import shard_batches
    assert shard_batches(7, 0, 2) == [0, 2, 4] and shard_batches(7, 1, 2) == [1, 3, 5]
    assert shard_batches(5, 0, 1) == [0, 1, 2, 3, 4]
    assert shard_batches(3, 3, 4) == []


def test_balanced_row_splits_balance_entries_not_rows():
    from glass_b200.partition import balanced_row_splits
    # 1000 rows: the first 10 hold half of all entries
    deg = torch.cat((torch.full((10,), 500), torch.full((990,), 5)))
    rowptr = torch.cat((torch.zeros(1, dtype=torch.int64), torch.cumsum(deg, 0))).to(torch.int32)
    for parts in (1, 2, 4, 8):
        b = balanced_row_splits(rowptr, parts)
        assert b[0] == 0 and b[-1] == 1000 and len(b) == parts + 1 and all(x <= y for x, y in zip(b, b[1:]))
        nnz = [int(rowptr[b[i + 1]] - rowptr[b[i]]) for i in range(parts)]
        assert max(nnz) <= int(rowptr[-1]) / parts + 500
    assert balanced_row_splits(rowptr, 2)[1] <= 11      # the 10 hub rows alone fill the first half
