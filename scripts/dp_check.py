"""torchrun --nproc-per-node P scripts/dp_check.py [workload] : label-batch DP, symmetric-memory exchange vs NCCL.

Checks (a) replicas stay bit-identical under the fused reduce-scatter + Adam + all-gather kernel, (b) its parameters
agree with the NCCL all-reduce + Adam path after the same steps, and times both as captured graphs."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from glass_b200 import run
from glass_b200.graphed import GraphedTrainStep

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
name = sys.argv[1] if len(sys.argv) > 1 else "em_user_shaped"
wl = bench.make_workload(name)
p, g = wl["params"], wl["g"]
x, ei, ew = g.x.to(dev), g.edge_index.to(dev), g.edge_attr.to(dev)
batches = [(a.to(dev), b.to(dev)) for a, b in bench.batches_for(wl, 40, rank, world)]
out = {"workload": name, "world": world}


def build(mode):
    os.environ["GLASS_B200_DP"] = mode
    torch.manual_seed(0)
    model = run.build_model(p["hidden_dim"], p["conv_layer"], 0.0, 1, p["pool"], p["z_ratio"], p["aggr"], wl["max_deg"],
                            wl["out_dim"], pretrained=wl["table"], device=dev)       # dropout 0: runs are comparable
    init = {k: v.clone() for k, v in model.state_dict().items()}
    step = GraphedTrainStep(model, wl["loss_fn"], x, ei, ew, batches[0][0], batches[0][1], p["lr"], warmup=2)
    step.reset_to(init)
    return model, step


results = {}
for mode in ("symm", "nccl"):
    model, step = build(mode)
    assert step.dp_mode == mode, (step.dp_mode, mode)
    losses = [float(step(*b)) for b in batches[:4]]                                   # eager steps
    sd = torch.cat([v.flatten().float() for v in model.state_dict().values()])
    gathered = [torch.empty_like(sd) for _ in range(world)]
    dist.all_gather(gathered, sd)
    same = all(torch.equal(gathered[0], t) for t in gathered)
    results[mode] = sd
    step.capture()
    for b in batches[4:10]:
        step(*b)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for b in batches[10:40]:
        step(*b)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 30], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sd2 = torch.cat([v.flatten().float() for v in model.state_dict().values()])
    gathered = [torch.empty_like(sd2) for _ in range(world)]
    dist.all_gather(gathered, sd2)
    same2 = all(torch.equal(gathered[0], t_) for t_ in gathered)
    err = int(step.opt.error.item()) if mode == "symm" else 0
    out[mode] = {"ms_per_step": float(t), "replicas_identical_eager": bool(same), "replicas_identical_graph": bool(same2),
                 "finite": bool(torch.isfinite(sd2).all()), "timeout_flag": err, "losses": losses}
    del step, model
diff = (results["symm"] - results["nccl"]).abs().max() / results["nccl"].abs().max()
out["symm_vs_nccl_rel_diff_after_4_steps"] = float(diff)
if rank == 0:
    print(json.dumps(out), flush=True)
torch.cuda.synchronize(); sys.stdout.flush(); os._exit(0)
