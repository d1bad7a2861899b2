"""GraphNorm fwd+bwd time (CUDA-graph replay, warm) over a size grid; run once per GLASS_B200_GN_FUSED_MAX
setting to compare the one-launch cluster kernel with the three-kernel path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glass_b200 import ops
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
print("GN_FUSED_MAX", os.environ.get("GLASS_B200_GN_FUSED_MAX", "default"))
for c in (8, 20, 64):
    for n in (1250, 2500, 5000, 10000, 20000):
        for p in (0.0, 0.3):
            x = torch.randn(n, c, device=dev, requires_grad=True)
            w, b, a = (torch.ones(c, device=dev, requires_grad=True) for _ in range(3))
            gout = torch.randn(n, c, device=dev)
            def step():
                out = ops.graph_norm(x, w, b, a, 1e-5, 2, p, True)
                out.backward(gout)
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                for _ in range(3): step()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(10): step()
            for _ in range(3): g.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): g.replay()
            e1.record(); torch.cuda.synchronize()
            print(f"c={c:3d} n={n:6d} p={p:.1f} elems={n*c:8d} launches={ops._lib.load().glass_graphnorm_launches(n, c)}  fwd+bwd {e0.elapsed_time(e1) * 1e3 / 200:7.2f} us")
