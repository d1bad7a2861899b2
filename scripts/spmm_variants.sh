#!/bin/bash
# time SpMM tuning variants at the em_user shape with CUDA events (L2 flushed between launches)
for v in ${VARIANTS:-0 1 2 3 4}; do
GLASS_SPMM_VARIANT=$v python - <<PY
import sys, os, torch
sys.path.insert(0, os.getcwd())
from glass_b200 import datasets, ops
torch.cuda.set_device(0)
g = datasets.load_dataset("em_user_shaped")
adj = ops.build_csr(g.edge_index.cuda(), g.edge_attr.cuda(), g.num_nodes, "gcn")
x = torch.randn(g.num_nodes, 64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3): ops.spmm(adj, x)
ts = []
for _ in range(10):
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ops.spmm(adj, x); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
print("variant $v: min %.1f us  avg %.1f us" % (min(ts), sum(ts) / len(ts)), flush=True)
PY
done
