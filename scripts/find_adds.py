"""Which Python lines launch aten::add / add_ / fill_ / zeros in one eager train step (torch.profiler stacks)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from glass_b200 import run, utils
from glass_b200.optim import FusedAdam
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
wl = bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else "em_user_shaped")
p, g = wl["params"], wl["g"]
model = run.build_model(p["hidden_dim"], p["conv_layer"], p["dropout"], 1, p["pool"], p["z_ratio"], p["aggr"], wl["max_deg"], wl["out_dim"], pretrained=wl["table"], device=dev).train()
x, ei, ew = g.x.to(dev), g.edge_index.to(dev), g.edge_attr.to(dev)
pos, y = [t.to(dev) for t in bench.batches_for(wl, 1, 0, 1)[0]]
opt = FusedAdam(model.parameters(), lr=1e-3)
def step():
    z = utils.MaxZOZ(x, pos)
    opt.zero_grad(set_to_none=True)
    loss = wl["loss_fn"](model(x, ei, ew, pos, z, id=0), y)
    loss.backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True, record_shapes=True) as prof:
    step(); torch.cuda.synchronize()
for e in prof.events():
    if e.name in ("aten::add", "aten::add_", "aten::fill_", "aten::zero_", "aten::mul", "aten::div", "aten::copy_") and e.device_time_total > 0:
        shapes = e.input_shapes
        stack = [s for s in (e.stack or []) if "glass_b200" in s or "bench" in s or "autograd" in s][:3]
        print(f"{e.name:12s} {e.device_time_total:7.1f} us shapes={shapes} thread={e.thread} stack={stack}")
