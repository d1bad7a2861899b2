"""GPU: a few launches of each tcgen05 pair-GEMM kernel at n = 57,333 (for ncu --set full --import-source on)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glass_b200 import _lib
import glass_b200.ops  # noqa: F401  (registers the ops)
torch.cuda.set_device(0); dev = "cuda:0"; n = 57333
_ops = torch.ops.glass_b200
lib = _lib.load()
P = _lib.GEMM_TCGEN05
for (k1, k2, h, act) in [(64, 0, 64, 2), (64, 64, 64, 0)]:
    a1 = torch.randn(n, k1, device=dev); a2 = torch.randn(n, k2, device=dev) if k2 else None
    k = k1 + k2
    w0 = torch.randn(h, k, device=dev); w1 = torch.randn(h, k, device=dev)
    b0 = torch.randn(h, device=dev); b1 = torch.randn(h, device=dev)
    mask = (torch.rand(n, device=dev) > 0.5).to(torch.uint8)
    out = torch.empty(n, h, device=dev); acts = torch.empty(n, 2 * h, device=dev) if act else None
    dout = torch.randn(n, h, device=dev)
    da1 = torch.empty(n, k1, device=dev); da2 = torch.empty(n, k2, device=dev) if k2 else None
    dw0, dw1 = torch.empty_like(w0), torch.empty_like(w1); db0, db1 = torch.empty_like(b0), torch.empty_like(b1)
    ws = torch.empty(lib.glass_pair_linear_mix_bwd_workspace_bytes(n, h, k), dtype=torch.uint8, device=dev)
    for _ in range(3):
        _ops.pair_linear_mix_fwd_(a1, a2, w0, b0, w1, b1, mask, 0.8, act, P, out, acts)
        _ops.pair_linear_mix_bwd_(dout, acts, a1, a2, w0, w1, mask, 0.8, act, P, da1, da2, dw0, db0, dw1, db1, ws)
torch.cuda.synchronize()
