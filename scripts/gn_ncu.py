"""GraphNorm fwd+bwd at the em_user shape, eager -- target of `ncu -k regex:"k_colsums|k_gn_"`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glass_b200 import ops
dev = torch.device("cuda", 0)
n, c, p = (int(sys.argv[1]), int(sys.argv[2])) + (float(sys.argv[3]),) if len(sys.argv) > 3 else (57333, 64, 0.5)
x = torch.randn(n, c, device=dev, requires_grad=True)
w, b, a = (torch.ones(c, device=dev, requires_grad=True) for _ in range(3))
gout = torch.randn(n, c, device=dev)
for _ in range(4):
    ops.graph_norm(x, w, b, a, 1e-5, 2, p, True).backward(gout)
torch.cuda.synchronize()
