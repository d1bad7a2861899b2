"""GPU: a few isolated SpMM launches at the em_user shape (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glass_b200 import datasets, ops
torch.cuda.set_device(0)
name = sys.argv[1] if len(sys.argv) > 1 else "em_user_shaped"
g = datasets.load_dataset(name)
adj = ops.build_csr(g.edge_index.cuda(), g.edge_attr.cuda(), g.num_nodes, "gcn")
x = torch.randn(g.num_nodes, 64, device="cuda")
for _ in range(4):
    y = ops.spmm(adj, x)
torch.cuda.synchronize()
print("nnz", adj.nnz, "n", adj.n)
