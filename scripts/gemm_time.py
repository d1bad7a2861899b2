"""GPU: warm timings (CUDA events, L2 flushed) of the tcgen05 pair-GEMM kernels at n = 57,333, H = 64."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glass_b200 import _lib, ops
torch.cuda.set_device(0); dev = "cuda:0"; n = 57333
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
_ops = torch.ops.glass_b200


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        flush.fill_(1)
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    return f"{sum(ms) / len(ms) * 1e3:7.1f} us (min {ms[0] * 1e3:6.1f})"


lib = _lib.load()
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
    P = _lib.GEMM_TCGEN05
    print(f"k1={k1} k2={k2} h={h} act={act}")
    print("  fwd            ", timeit(lambda: _ops.pair_linear_mix_fwd_(a1, a2, w0, b0, w1, b1, mask, 0.8, act, P, out, acts)))
    print("  fwd (no acts)  ", timeit(lambda: _ops.pair_linear_mix_fwd_(a1, a2, w0, b0, w1, b1, mask, 0.8, act, P, out, None)))
    print("  bwd dX+dW      ", timeit(lambda: _ops.pair_linear_mix_bwd_(dout, acts, a1, a2, w0, w1, mask, 0.8, act, P, da1, da2, dw0, db0, dw1, db1, ws)))
    print("  bwd dW only    ", timeit(lambda: _ops.pair_linear_mix_bwd_(dout, acts, a1, a2, w0, w1, mask, 0.8, act, P, None, None, dw0, db0, dw1, db1, ws)))
    if k2:
        stats = torch.zeros(6, k1, device=dev); stats[0] = 1.0
        bits = torch.randint(-2**31, 2**31 - 1, ((n * k1 + 31) // 32,), device=dev, dtype=torch.int32)
        print("  fwd normalised ", timeit(lambda: _ops.pair_linear_mix_fwd_ex_(a1, a2, w0, b0, w1, b1, mask, 0.8, act, P, out, None, stats, bits, 0.5, 0, None, None, 0.0, 0)))
        print("  bwd normalised ", timeit(lambda: _ops.pair_linear_mix_bwd_ex_(dout, acts, a1, a2, w0, w1, mask, 0.8, act, P, da1, da2, dw0, db0, dw1, db1, ws, stats, bits, 0.5, 0, None, None, 0.0, 0, 0, 0)))
