"""Warm, in-situ kernel times of graph-replayed train steps (torch.profiler / CUPTI), em_user shape."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from glass_b200 import run
from glass_b200.graphed import GraphedTrainStep
torch.cuda.set_device(0); dev = torch.device("cuda", 0)
wl = bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else "em_user_shaped")
p, g = wl["params"], wl["g"]
torch.manual_seed(0)
model = run.build_model(p["hidden_dim"], p["conv_layer"], p["dropout"], 1, p["pool"], p["z_ratio"], p["aggr"], wl["max_deg"], wl["out_dim"], pretrained=wl["table"], device=dev)
x, ei, ew = g.x.to(dev), g.edge_index.to(dev), g.edge_attr.to(dev)
batches = [(a.to(dev), b.to(dev)) for a, b in bench.batches_for(wl, 30, 0, 1)]
step = GraphedTrainStep(model, wl["loss_fn"], x, ei, ew, batches[0][0], batches[0][1], p["lr"]).capture()
for a, b in batches[:10]: step(a, b)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for a, b in batches[10:30]: step(a, b)
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        import re
        m = re.search(r"(k_[A-Za-z0-9_]+(?:<[^>]*>)?)", e.name)
        if m:
            name = m.group(1)[:40]
        else:   # torch kernels: keep the functor, it tells which op launched it
            f = re.search(r"(\w+Functor\w*|\w+_kernel_cuda\w*|\w+Op<|\w+Ops<|\w+Impl\b)", e.name)
            name = (e.name[:24] + ".." + f.group(1)[:28]) if f else e.name[:56]
        a_ = agg.setdefault(name, [0, 0.0]); a_[0] += 1; a_[1] += e.device_time
tot = sum(v[1] for v in agg.values())
print(f"20 graph replays: kernel time {tot/20:.1f} us/step")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:44]:
    print(f"{t/20:8.1f} us/step  x{c/20:4.1f}  {k}")
