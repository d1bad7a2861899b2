"""GPU: isolated k_spmm timings (L2 flushed, CUDA events) next to the L2 gather probe (glass_l2_gather_probe: the same
256-byte row gathers with no index stream and no per-row FMA chain) on the named graphs."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glass_b200 import _lib, datasets, ops

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


out = {}
for name in sys.argv[1:] or ["em_user_shaped", "em_user_shaped_powerlaw"]:
    g = datasets.load_dataset(name, device=dev) if name.startswith("stress") else datasets.load_dataset(name)
    adj = ops.build_csr(g.edge_index.to(dev), g.edge_attr.to(dev), g.num_nodes, "gcn")
    n, h = adj.n, 64
    x = torch.randn(n, h, device=dev)
    y = torch.empty(n, h, device=dev)
    sink = torch.empty(lib.glass_sm_count() * 5 * 256, device=dev)
    probe = lambda: _lib.check(lib.glass_l2_gather_probe(C.c_void_p(x.data_ptr()), x.stride(0), n, h, adj.nnz,
                                                         C.c_void_p(sink.data_ptr()), sink.numel(), None), "probe")
    tp = timeit(probe)
    ts = timeit(lambda: ops._run_spmm(adj.rowptr, adj.col, adj.val, adj.plan, x, y))
    tt = timeit(lambda: ops._run_spmm(adj.rowptr_t, adj.col_t, adj.val_t, adj.plan_t, x, y))
    gb = adj.nnz * 4 * h
    out[name] = {"n": n, "nnz": adj.nnz, "planned": adj.plan is not None, "probe_us": tp, "spmm_us": ts, "spmm_t_us": tt,
                 "probe_gather_GBps": round(gb / tp[0] * 1e-3, 1), "spmm_gather_GBps": round(gb / ts[0] * 1e-3, 1),
                 "spmm_frac_of_probe": round(tp[0] / ts[0], 3)}
    print(name, json.dumps(out[name]), flush=True)
