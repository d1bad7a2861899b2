"""GPU: isolated k_spmm timings (L2 flushed before each launch, CUDA events) over the dispatcher's tuning variants,
grid caps, the statistics epilogue and the row-split length, on the em_user-shaped graphs."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glass_b200 import _lib, datasets, ops

torch.cuda.set_device(0)
lib = _lib.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, reps=12):
    for _ in range(3):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        flush.fill_(1)
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    return round(sum(ms) / len(ms) * 1e3, 1), round(ms[0] * 1e3, 1)


out = {}
for name in sys.argv[1:] or ["em_user_shaped", "em_user_shaped_powerlaw"]:
    g = datasets.load_dataset(name)
    adj = ops.build_csr(g.edge_index.cuda(), g.edge_attr.cuda(), g.num_nodes, "gcn")
    n, h = adj.n, 64
    x = torch.randn(n, h, device="cuda")
    y = torch.empty(n, h, device="cuda")
    partial = ops._stats_table(h, x.device)
    res = out[name] = {}
    split_lens = [None] if adj.plan is None else [512, 256, 128, 64]
    for sl in split_lens:
        if sl is not None:
            adj.make_plans(sl)
        for variant in range(8):
            for waves in ((0,) if variant else (0, 1, 2, 4)):
                for stats in (False, True):
                    lib.glass_tune(b"spmm_variant", variant)
                    lib.glass_tune(b"spmm_waves", waves)
                    key = f"split{sl}_v{variant}_w{waves}_{'stats' if stats else 'plain'}"
                    try:
                        res[key] = timeit(lambda: ops._run_spmm(adj.rowptr, adj.col, adj.val, adj.plan, x, y,
                                                                partial if stats else None))
                    except Exception as e:  # noqa: BLE001
                        res[key] = str(e)[:100]
                    print(name, key, res[key], flush=True)
    lib.glass_tune(b"spmm_variant", 0)
    lib.glass_tune(b"spmm_waves", 0)
    # statistics parity: epilogue sums vs fp64 column sums of y
    nblk = ops._run_spmm(adj.rowptr, adj.col, adj.val, adj.plan, x, y, partial)
    s = partial[:h, :nblk].sum(1)
    q = partial[h:, :nblk].sum(1)
    yd = y.double()
    res["stats_check"] = [float((s - yd.sum(0)).abs().max() / yd.sum(0).abs().max()),
                          float((q - (yd * yd).sum(0)).abs().max() / (yd * yd).sum(0).abs().max()), nblk]
    print(name, "stats_check", res["stats_check"], flush=True)
json.dump(out, open("gpurun_out/spmm_sweep.json", "w"), indent=1)
