"""GPU diagnostic: tcgen05 pair-GEMM (fwd and dX) against the SIMT fp32 kernels, with error structure."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glass_b200 import _lib, ops

torch.cuda.set_device(0)
dev = "cuda:0"


def run(n, k1, k2, h, act, z):
    g = torch.Generator().manual_seed(n + h)
    a1 = torch.randn(n, k1, generator=g).to(dev)
    a2 = torch.randn(n, k2, generator=g).to(dev) if k2 else None
    k = k1 + k2
    w0, w1 = (torch.randn(h, k, generator=g) / k ** 0.5).to(dev), (torch.randn(h, k, generator=g) / k ** 0.5).to(dev)
    b0, b1 = torch.randn(h, generator=g).to(dev), torch.randn(h, generator=g).to(dev)
    mask = (torch.rand(n, generator=g) > 0.5).to(torch.uint8).to(dev)
    gout = torch.randn(n, h, generator=g).to(dev)
    res = {}
    for name, pid in (("simt", _lib.GEMM_SIMT), ("tc", _lib.GEMM_TCGEN05)):
        t = [x.clone().requires_grad_(True) if x is not None else None for x in (a1, a2, w0, b0, w1, b1)]
        out = ops.pair_linear_mix(t[0], t[1], t[2], t[3], t[4], t[5], mask, z, act, pid)
        out.backward(gout)
        torch.cuda.synchronize()
        res[name] = [out.detach()] + [x.grad for x in t if x is not None]
    names = ["out", "da1"] + (["da2"] if k2 else []) + ["dw0", "db0", "dw1", "db1"]
    line = f"n={n} k1={k1} k2={k2} h={h} act={act}: "
    for nm, s, t in zip(names, res["simt"], res["tc"]):
        err = float((s - t).abs().max() / s.abs().max().clamp(min=1e-30))
        line += f"{nm}={err:.1e} "
        if err > 1e-4 and nm in ("out", "da1", "da2"):
            bad = ((s - t).abs() > 1e-3 * s.abs().max()).nonzero()
            rows = torch.unique(bad[:, 0])[:12].tolist()
            cols = torch.unique(bad[:, 1])[:16].tolist()
            line += f"\n   BAD {nm}: {bad.shape[0]} of {s.numel()} entries; rows {rows} cols {cols}; sample simt {s[bad[0,0], bad[0,1]].item():.4f} tc {t[bad[0,0], bad[0,1]].item():.4f}\n   "
    print(line, flush=True)


for cfg in [(128, 32, 0, 8, 0, 0.8), (128, 32, 0, 64, 0, 0.8), (128, 64, 0, 64, 2, 0.8), (1000, 64, 0, 64, 2, 0.8),
            (777, 64, 64, 64, 0, 0.75), (4096, 128, 0, 128, 2, 0.6), (57333, 64, 64, 64, 0, 0.75), (130, 64, 64, 32, 2, 0.8),
            (513, 8, 8, 8, 0, 1.0), (300, 16, 0, 16, 1, 0.5)]:
    try:
        run(*cfg)
    except Exception as e:
        print(cfg, "EXC", type(e).__name__, e, flush=True)

# timing at the em_user shape
for (n, k1, k2, h, act) in [(57333, 64, 0, 64, 2), (57333, 64, 64, 64, 0)]:
    g = torch.Generator().manual_seed(0)
    a1 = torch.randn(n, k1, generator=g).to(dev); a2 = torch.randn(n, k2, generator=g).to(dev) if k2 else None
    k = k1 + k2
    w0 = torch.randn(h, k).to(dev); w1 = torch.randn(h, k).to(dev); b0 = torch.randn(h).to(dev); b1 = torch.randn(h).to(dev)
    mask = torch.ones(n, dtype=torch.uint8, device=dev)
    for name, pid in (("simt", _lib.GEMM_SIMT), ("tc", _lib.GEMM_TCGEN05)):
        try:
            for _ in range(3):
                ops.pair_linear_mix(a1, a2, w0, b0, w1, b1, mask, 0.8, act, pid)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                ops.pair_linear_mix(a1, a2, w0, b0, w1, b1, mask, 0.8, act, pid)
            e1.record(); torch.cuda.synchronize()
            print(f"fwd n={n} k={k} h={h} {name}: {e0.elapsed_time(e1)/20*1e3:.1f} us", flush=True)
        except Exception as e:
            print(name, "EXC", e)
