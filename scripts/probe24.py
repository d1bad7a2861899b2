"""GPU: L2 gather probe on fp32 rows next to the 24-bit two-plane copy (glass_l2_gather_probe24), em_user shape."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glass_b200 import _lib
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
lib = _lib.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        flush.fill_(1)
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    return round(sum(ms) / len(ms) * 1e3, 1), round(ms[0] * 1e3, 1)


n, h, gathers = 57333, 64, 9146834
x = torch.randn(n, h, device=dev)
hi = torch.empty(n, h, dtype=torch.int16, device=dev)
lo = torch.empty(n, h, dtype=torch.uint8, device=dev)
sink = torch.empty(lib.glass_sm_count() * 5 * 256, device=dev)
vp = lambda t: C.c_void_p(t.data_ptr())
p32 = lambda: _lib.check(lib.glass_l2_gather_probe(vp(x), x.stride(0), n, h, gathers, vp(sink), sink.numel(), None), "probe")
s32 = None
p24 = lambda conv=0: _lib.check(lib.glass_l2_gather_probe24(vp(x), n, h, gathers, vp(hi), vp(lo), vp(sink), sink.numel(), conv, None), "probe24")
p32(); ref = sink.clone()
p24(1); got = sink.clone()
torch.cuda.synchronize()
rel = float((got.double() - ref.double()).abs().max() / ref.double().abs().max())
out = {"fp32_us": timeit(p32), "split24_us": timeit(lambda: p24(0)), "split24_with_conversion_us": timeit(lambda: p24(1)),
       "sum_rel_diff": rel, "bytes_fp32": gathers * 4 * h, "bytes_24": gathers * 3 * h}
for u in (4, 8):
    p256 = lambda u=u: _lib.check(lib.glass_l2_gather_probe256(vp(x), x.stride(0), n, h, gathers, vp(sink), sink.numel(), u, None), "probe256")
    out[f"fp32_256bit_loads_unroll{u}_us"] = timeit(p256)
print(json.dumps(out))
