"""torchrun --nproc-per-node P scripts/partition_check.py [graph] : the row-partitioned GLASS model on
P GPUs against the replicated single-GPU model on the same graph / weights / label batch: logits, every parameter
gradient and the gradient of the input-embedding shard; then timings of one forward+backward."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from glass_b200 import datasets, ops, run, utils
from glass_b200.partition import PartitionedGLASS, RowPartitionedAdj

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
name = sys.argv[1] if len(sys.argv) > 1 else "stress_small"
g = datasets.load_dataset(name, device=dev)
n, h = g.num_nodes, 64
p = run.load_params(name)
ei, ew = g.edge_index.to(dev), g.edge_attr.to(dev)
torch.manual_seed(0)
table = datasets.synthetic_embedding(n, h, 0)
model = run.build_model(h, 1, 0.0, 1, p["pool"], p["z_ratio"], p["aggr"], n - 1, 1, pretrained=table, device=dev).train()
pos = g.pos[:8].to(dev)
y = g.y[:8].to(dev).float()
x = torch.arange(n, device=dev).reshape(n, 1, 1)
loss_fn = torch.nn.BCEWithLogitsLoss()
z = utils.MaxZOZ(x, pos)

# replicated reference on this GPU
model.zero_grad(set_to_none=True)
ref_logits = model(x, ei, ew, pos, z)
loss_fn(ref_logits.flatten(), y).backward()
ref = {k: v.grad.detach().clone() for k, v in model.named_parameters()}

adj = model.conv.convs[0].adj
pipelined = os.environ.get('PIPELINED', '0') != '0'
part = RowPartitionedAdj(adj, rank, world, pipelined=pipelined)
pm = PartitionedGLASS(model, part)
model.zero_grad(set_to_none=True)
h_local = model.conv.input_emb.weight.detach()[part.lo:part.hi].clone().requires_grad_(True)
logits = pm(h_local, pos, z)
loss_fn(logits.flatten(), y).backward()
pm.reduce_grads()


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp(min=1e-30))


errs = {"logits": rel(logits, ref_logits), "table_shard": rel(h_local.grad, ref["conv.input_emb.weight"][part.lo:part.hi])}
for k, v in model.named_parameters():
    if k != "conv.input_emb.weight" and v.grad is not None:
        errs[k] = rel(v.grad, ref[k])
worst = torch.tensor([max(errs.values())], device=dev)
if world > 1:
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def step_part():
    model.zero_grad(set_to_none=True)
    h_local.grad = None
    loss_fn(pm(h_local, pos, z).flatten(), y).backward()
    pm.reduce_grads()


def step_repl():
    model.zero_grad(set_to_none=True)
    loss_fn(model(x, ei, ew, pos, z).flatten(), y).backward()


out = {"graph": name, "world": world, "pipelined": pipelined, "n": n, "nnz": int(adj.nnz), "worst_rel_err": float(worst),
       "ms_partitioned_fwd_bwd": timed(step_part), "ms_replicated_fwd_bwd": timed(step_repl),
       "nnz_local": part.nnz_local, "errs_rank0": {k: round(v, 9) for k, v in sorted(errs.items(), key=lambda kv: -kv[1])[:6]}}
if rank == 0:
    print(json.dumps(out), flush=True)
torch.cuda.synchronize(); sys.stdout.flush(); os._exit(0)
