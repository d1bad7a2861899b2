"""GPU: the fused last-GraphNorm + pooling operator (forward + backward) at the em_user shape, eager -- target of
`ncu -k regex:"k_pool_pad|k_colsums|k_gn_finalize|k_mark_nodes"`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from glass_b200 import ops
dev = torch.device("cuda", 0)
wl = bench.make_workload("em_user_shaped")
pos = bench.batches_for(wl, 1, 0, 1)[0][0].to(dev)
n, c = wl["g"].num_nodes, 64
x = torch.randn(n, c, device=dev, requires_grad=True)
w, b, a = (torch.ones(c, device=dev, requires_grad=True) for _ in range(3))
for _ in range(4):
    out = ops.graph_norm_pool(x, w, b, a, 1e-5, pos, "size")
    out.backward(torch.randn_like(out))
torch.cuda.synchronize()
