"""torchrun --nproc-per-node P scripts/profile_step_dp.py [workload]: warm, in-situ kernel times of graph-replayed
data-parallel train steps (torch.profiler / CUPTI) on rank 0, plus the device time per step."""
import sys, os, collections, re
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from glass_b200 import run
from glass_b200.graphed import GraphedTrainStep
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
wl = bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else "em_user_shaped")
p, g = wl["params"], wl["g"]
torch.manual_seed(0)
model = run.build_model(p["hidden_dim"], p["conv_layer"], p["dropout"], 1, p["pool"], p["z_ratio"], p["aggr"], wl["max_deg"], wl["out_dim"], pretrained=wl["table"], device=dev)
x, ei, ew = g.x.to(dev), g.edge_index.to(dev), g.edge_attr.to(dev)
batches = [(a.to(dev), b.to(dev)) for a, b in bench.batches_for(wl, 60, rank, world)]
step = GraphedTrainStep(model, wl["loss_fn"], x, ei, ew, batches[0][0], batches[0][1], p["lr"]).capture()
for a, b in batches[:10]: step(a, b)
torch.cuda.synchronize()
if world > 1: dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for a, b in batches[10:40]: step(a, b)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 30
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for a, b in batches[40:60]: step(a, b)
    torch.cuda.synchronize()
if rank == 0:
    agg = collections.OrderedDict()
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            m = re.search(r"(k_[A-Za-z0-9_]+(?:<[^>]*>)?)", e.name)
            if m:
                name = m.group(1)[:40]
            else:
                f = re.search(r"(\w+Functor\w*|\w+_kernel_cuda\w*|\w+Op<|\w+Ops<|\w+Impl\b)", e.name)
                name = (e.name[:24] + ".." + f.group(1)[:28]) if f else e.name[:56]
            a_ = agg.setdefault(name, [0, 0.0]); a_[0] += 1; a_[1] += e.device_time
    tot = sum(v[1] for v in agg.values())
    print(f"world {world} dp_mode {step.dp_mode}: {ms * 1e3:.1f} us/step device time; kernel time {tot / 20:.1f} us/step")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"{t / 20:8.1f} us/step  x{c / 20:4.1f}  {k}")
torch.cuda.synchronize(); sys.stdout.flush(); os._exit(0)
