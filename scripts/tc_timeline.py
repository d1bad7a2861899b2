import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glass_b200 import _lib, ops
torch.cuda.set_device(0); dev = "cuda:0"; n = 57333
for (k1, k2, h, act) in [(64, 0, 64, 2), (64, 64, 64, 0)]:
    a1 = torch.randn(n, k1, device=dev).requires_grad_(True)
    a2 = torch.randn(n, k2, device=dev).requires_grad_(True) if k2 else None
    k = k1 + k2
    w0 = torch.randn(h, k, device=dev).requires_grad_(True); w1 = torch.randn(h, k, device=dev).requires_grad_(True)
    b0 = torch.randn(h, device=dev).requires_grad_(True); b1 = torch.randn(h, device=dev).requires_grad_(True)
    mask = (torch.rand(n, device=dev) > 0.5).to(torch.uint8)
    for _ in range(2):
        out = ops.pair_linear_mix(a1, a2, w0, b0, w1, b1, mask, 0.8, act, _lib.GEMM_TCGEN05)
        out.backward(torch.randn(n, h, device=dev))
    torch.cuda.synchronize()
