#!/usr/bin/env python
"""Benchmark of the GLASS labeled message-passing hot path on B200 (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload em_user_shaped] [--impl reference]

A "step" is one training iteration of impl/train.py:10-16 (labels z, forward, loss, backward, optimizer)
over one label batch on the whole base graph.  Default workload: the em_user-shaped synthetic graph with
the reference's em_user hyper-parameters (BASELINE.json configs[3], the roofline config).  Prints ONE JSON line
(contract in the task statement): device-timed `value`, host-buffer `e2e`, the SpMM `roofline`, and the
CPU `cpu_baseline` (oracle port of the reference modules, timed on this box's host cores).
`--impl reference` times that CPU port alone with the same metric/config keys.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHIPPED = ("density", "cut_ratio", "coreness", "component")


# ------------------------------------------------------------------------------------------ workload
def make_workload(name: str, seed: int = 0):
    """Graph, hyper-parameters, node ids, (synthetic) embedding table and the train split, on the host."""
    from glass_b200 import datasets, run
    torch.manual_seed(seed)
    g = datasets.load_dataset(name)
    params = run.load_params(name)
    loss_fn, out_dim, score_fn, y = run.task_of(g.y)
    g.y = y
    if name in SHIPPED:
        g.setOneFeature()                      # --use_one (README.md:50-53 commands for synthetic sets)
        table = None
    else:
        g.setNodeIdFeature()                   # --use_nodeid with a seeded stand-in for Emb/<name>_64.pt
        table = datasets.synthetic_embedding(g.num_nodes, params["hidden_dim"], seed)
    trn = g.get_split("train")
    return dict(name=name, g=g, params=params, loss_fn=loss_fn, out_dim=out_dim, table=table,
                trn_pos=trn[3], trn_y=trn[4], max_deg=int(g.x.max()))


def batches_for(wl, n_steps: int, rank: int, world: int, seed: int = 0):
    """Host (pos, y) batches: one shared seeded order, rank r takes batches r, r+P, ... of every epoch."""
    from glass_b200.dist import shard_batches, shared_permutation
    bs = wl["params"]["batch_size"]
    n_trn = wl["trn_pos"].shape[0]
    per_epoch = n_trn // bs                    # drop_last=True like GLASSTest.py:106-116
    out, epoch = [], 0
    while len(out) < n_steps:
        perm = shared_permutation(n_trn, seed, epoch)
        for b in shard_batches(per_epoch, rank, world):
            idx = perm[b * bs:(b + 1) * bs]
            out.append((wl["trn_pos"][idx].contiguous(), wl["trn_y"][idx].contiguous()))
            if len(out) == n_steps:
                break
        epoch += 1
    return out


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU during the timed region (pynvml)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.stop = index, [], set(), False
        self.max_mhz, self.thread = None, None

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)

            def loop():
                while not self.stop:
                    try:
                        self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                        for bit, nm in self.REASONS.items():
                            if mask & bit:
                                self.reasons.add(nm)
                    except Exception:
                        pass
                    time.sleep(0.02)

            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
        except Exception:
            pass
        return self

    def __exit__(self, *a):
        self.stop = True
        if self.thread:
            self.thread.join(timeout=1)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------ CPU port (reference arm)
def cpu_port_steps(wl, n_steps: int, warmup: int, train: bool = True, device=None):
    """Times the oracle port of the reference modules (same ATen op sequence as impl/models.py).  On the CPU
    this is the `cpu_baseline` / `--impl reference` arm; with a CUDA `device` it is the reference's own eager
    torch.sparse path on the GPU (reported only as an extra, `--gpu-eager-baseline`)."""
    from oracle import glass_oracle as O
    p = wl["params"]
    g = wl["g"]
    cfg = O.GlassConfig(hidden_dim=p["hidden_dim"], conv_layer=p["conv_layer"], aggr=p["aggr"], z_ratio=p["z_ratio"],
                        dropout=p["dropout"], pool=p["pool"], jk=True, activation="elu", out_dim=wl["out_dim"])
    sd = O.init_state_dict(cfg, wl["max_deg"] + 1, seed=0, pretrained=wl["table"])
    if device is not None:
        sd = {k: v.to(device) for k, v in sd.items()}
    model = O.OracleModel(cfg, sd)
    opt = torch.optim.Adam(model.params, lr=p["lr"])
    loss_fn = O.loss_fn_for(wl["out_dim"] == 1 and g.y.dtype == torch.float32)
    mv = (lambda t: t.to(device)) if device is not None else (lambda t: t)
    gx, gei, gea = mv(g.x), mv(g.edge_index), mv(g.edge_attr)
    batches = [(mv(a), mv(b)) for a, b in batches_for(wl, n_steps + warmup, 0, 1)]
    for pos, y in batches[:warmup]:
        model.step(opt, gx, gei, gea, pos, y, loss_fn, training=train)
    if device is not None:
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for pos, y in batches[warmup:]:
        model.step(opt, gx, gei, gea, pos, y, loss_fn, training=train)
    if device is not None:
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return p["batch_size"] * n_steps / dt, dt / n_steps


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    wl = make_workload(args.workload)
    value, per_step = cpu_port_steps(wl, args.steps, args.warmup)
    cores = torch.get_num_threads()
    line = base_line(args, wl, value, per_step * 1e3)
    line.update(impl="reference", n_gpus=args.gpus, dtype="f32",
                cpu_baseline={"value": value, "unit": "subgraphs/s", "cores": cores, "kind": "port",
                              "sample": f"{args.steps} train steps of batch {wl['params']['batch_size']} on the full "
                                        f"{wl['name']} graph (oracle port of impl/models.py; /root/reference is not on this box)"},
                e2e={"value": value, "unit": "subgraphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                gpu_launches=0)
    print(json.dumps(line), flush=True)


def base_line(args, wl, value, ms_per_step):
    p = wl["params"]
    g = wl["g"]
    return {"metric": "subgraphs/s (train)", "value": value, "unit": "subgraphs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "nodes": g.num_nodes, "nnz": int(g.edge_index.shape[1]),
                       "hidden_dim": p["hidden_dim"], "conv_layer": p["conv_layer"], "aggr": p["aggr"],
                       "pool": p["pool"], "batch_size_per_gpu": p["batch_size"], "dropout": p["dropout"],
                       "z_ratio": p["z_ratio"], "subgraph_pad": int(wl["trn_pos"].shape[1]),
                       "parallelism": f"label-batch dp{args.gpus}",
                       "l2": "flushed (256 MiB write) between isolated SpMM timings; whole-step working set "
                             "(A, A^T, activations) exceeds L2"}}


# ------------------------------------------------------------------------------------------ product arm
def spmm_roofline(adj, h: int, reps: int = 20, workload: str = None):
    """Isolated SpMM launches, L2 flushed before each, CUDA events on the launching stream."""
    from glass_b200 import ops
    dev = adj.col.device
    x = torch.randn(adj.n, h, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for _ in range(3):
        ops.spmm(adj, x)
    for a, b in ev:
        flush.fill_(1)
        a.record()
        ops.spmm(adj, x)
        b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    avg = sum(ms) / len(ms)
    algo = 4 * (adj.n + 1) + 8 * adj.nnz + 8 * adj.n * h
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = algo / (avg * 1e-3) / 1e9
    # DRAM bytes per launch come from an `ncu --set full` capture (profiles/); they are only reported when the
    # capture was taken from THIS build of spmm.cu (digest of the source stamped into the file) and this graph
    traffic, traffic_note = None, "no ncu capture for this build of csrc/spmm.cu"
    try:
        import hashlib
        with open(os.path.join(ROOT, "glass_b200", "csrc", "spmm.cu"), "rb") as f:
            digest = hashlib.sha256(f.read()).hexdigest()[:16]
        with open(os.path.join(ROOT, "profiles", "spmm_traffic.json")) as f:
            for t in json.load(f)["captures"]:
                if (t.get("nnz") == adj.nnz and t.get("h") == h and t.get("spmm_cu_sha16") == digest and
                        (workload is None or t.get("workload") == workload)):
                    traffic, traffic_note = t["dram_bytes_per_launch"], t.get("source", "")
    except Exception:
        pass
    # second roofline: the kernel's real bound is the L2 -> SM gather stream (4 H nnz bytes; x is L2 resident).  The
    # probe reads the same number of pseudo-random 4H-byte rows with the same lane layout, 8 loads in flight, no index
    # stream and no dependent FMA chain: the gather rate this device can deliver to ANY kernel of this shape.
    l2 = None
    if h in (32, 64, 128) and adj.n * h < 2 ** 31:
        try:      # a secondary number: it must not take the headline line down
            import ctypes as C
            from glass_b200 import _lib
            lib = _lib.load()
            sink = torch.empty(lib.glass_sm_count() * 5 * 256, device=dev)
            probe = lambda: _lib.check(lib.glass_l2_gather_probe(C.c_void_p(x.data_ptr()), x.stride(0), adj.n, h, adj.nnz,
                                                                 C.c_void_p(sink.data_ptr()), sink.numel(),
                                                                 C.c_void_p(torch.cuda.current_stream().cuda_stream)), "probe")
            for _ in range(3):
                probe()
            for a, b in ev:
                flush.fill_(1)
                a.record()
                probe()
                b.record()
            torch.cuda.synchronize()
            pm = sum(a.elapsed_time(b) for a, b in ev) / len(ev)
            gb = 4 * h * adj.nnz
            l2 = {"probe_gbs": gb / (pm * 1e-3) / 1e9, "probe_us": pm * 1e3, "spmm_gbs": gb / (avg * 1e-3) / 1e9,
                  "frac": pm / avg, "what": "glass_l2_gather_probe: nnz pseudo-random 4H-byte row gathers from the same x "
                                            "(L2 flushed before each launch); frac = k_spmm gather rate / probe gather rate"}
        except Exception as e:  # noqa: BLE001
            l2 = {"error": f"{type(e).__name__}: {e}"[:200]}
    return {"bound": "hbm", "kernel": "k_spmm (glass_spmm_csr)", "achieved": achieved, "peak": peak, "unit": "GB/s", "l2_gather": l2,
            "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_note, "algorithmic_bytes": algo, "us_per_launch": avg * 1e3,
            "us_min": ms[0] * 1e3, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
            "gather_bytes_l2": 4 * h * adj.nnz, "gather_gbs_l2": 4 * h * adj.nnz / (avg * 1e-3) / 1e9,
            "peak_nominal": 8000.0, "frac_nominal": achieved / 8000.0}


def kernel_rooflines(model, wl, x, ei, ew, pos, dev, peak_gbs: float):
    """Isolated timings (L2 flushed before every launch, CUDA events on the launching stream) of the other kernels of
    the path with their algorithmic bytes: achieved GB/s and the fraction of the measured HBM peak."""
    from glass_b200 import _lib, ops
    from glass_b200.optim import FusedAdam
    p = wl["params"]
    n, h = int(x.shape[0]), p["hidden_dim"]
    nnz = int(ei.shape[1])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    _ops = torch.ops.glass_b200

    def timed(fn, reps=8):
        for _ in range(2):
            fn()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for a, b in ev:
            flush.fill_(1)
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in ev) / reps * 1e3

    out = {}

    def add(name, us, nbytes, launches):
        out[name] = {"us": round(us, 2), "algorithmic_bytes": int(nbytes), "gbs": round(nbytes / us * 1e-3, 1),
                     "frac_hbm": round(nbytes / us * 1e-3 / peak_gbs, 4), "launches": launches}

    # csr_build: int64 pairs + weights read, CSR + transposed CSR written (init path, one host sync inside)
    add("csr_build", timed(lambda: ops.build_csr(ei, ew, n, p["aggr"]), reps=3), 20 * nnz + 2 * (8 * nnz + 4 * (n + 1)), 9)
    f32 = dict(dtype=torch.float32, device=dev)
    a = torch.randn(n, h, **f32)
    g = torch.randn(n, h, **f32)
    w, b, ms = torch.ones(h, **f32), torch.zeros(h, **f32), torch.ones(h, **f32)
    mask = torch.zeros(n, dtype=torch.uint8, device=dev)
    # GraphNorm forward / backward (one cooperative launch each at this size): x read once, out written once
    ar = a.clone().requires_grad_(True)
    add("graphnorm_fwd", timed(lambda: ops.graph_norm(a, w, b, ms, 1e-5, 0, 0.5, True)), 8 * n * h,
        _lib.load().glass_graphnorm_launches(n, h))
    y = ops.graph_norm(ar, w, b, ms, 1e-5, 0, 0.5, True)
    add("graphnorm_bwd", timed(lambda: y.backward(g, retain_graph=True)), 12 * n * h,
        _lib.load().glass_graphnorm_launches(n, h))
    # label-mixed Linear pairs (tcgen05): operands read once, out (+ saved activations) written once
    w0, w1 = torch.randn(h, h, **f32) / 8, torch.randn(h, h, **f32) / 8
    c0, c1 = torch.randn(h, 2 * h, **f32) / 11, torch.randn(h, 2 * h, **f32) / 11
    b0, b1 = torch.zeros(h, **f32), torch.zeros(h, **f32)
    o = torch.empty(n, h, **f32)
    acts = torch.empty(n, 2 * h, **f32)
    path = _lib.GEMM_AUTO
    add("pair_trans_fwd", timed(lambda: _ops.pair_linear_mix_fwd_(a, None, w0, b0, w1, b1, mask, 0.75, 2, path, o, acts)),
        4 * n * (h + h + 2 * h), 1)
    add("pair_comb_fwd", timed(lambda: _ops.pair_linear_mix_fwd_(a, g, c0, b0, c1, b1, mask, 0.75, 0, path, o, None)),
        4 * n * (2 * h + h), 1)
    ws = torch.empty(_lib.load().glass_pair_linear_mix_bwd_workspace_bytes(n, h, 2 * h), dtype=torch.uint8, device=dev)
    da1, da2 = torch.empty(n, h, **f32), torch.empty(n, h, **f32)
    dw0, dw1, db0, db1 = torch.empty_like(c0), torch.empty_like(c1), torch.empty(h, **f32), torch.empty(h, **f32)
    add("pair_comb_bwd", timed(lambda: _ops.pair_linear_mix_bwd_(o, None, a, g, c0, c1, mask, 0.75, 0, path, da1, da2, dw0,
                                                                db0, dw1, db1, ws)), 4 * n * (h + 2 * h + 2 * h) + 4 * n * (h + 2 * h), 3)
    # pooling over one label batch: ids + gathered rows + output forward; the ordered backward writes every row of demb
    d = h * p["conv_layer"]
    emb = torch.randn(n, d, **f32).requires_grad_(True)
    n_valid = int((pos >= 0).sum())
    add("segment_pool_fwd", timed(lambda: ops.segment_pool(emb, pos, p["pool"])), 8 * pos.numel() + 4 * d * n_valid + 4 * pos.shape[0] * d, 1)
    pooled = ops.segment_pool(emb, pos, p["pool"])
    gp = torch.randn_like(pooled)
    add("segment_pool_bwd", timed(lambda: pooled.backward(gp, retain_graph=True)), 8 * pos.numel() + 4 * n * d + 4 * pos.shape[0] * d, 3)
    # multi-tensor Adam over all parameters: p, m, v read and written, g read (7 passes)
    params = [q for q in model.parameters() if q.requires_grad]
    for q in params:
        q.grad = torch.zeros_like(q)
    n_par = sum(q.numel() for q in params)
    opt = FusedAdam(params, lr=0.0)
    add("adam", timed(opt.step), 28 * n_par, 1)
    for q in params:
        q.grad = None
    return out


def run_product(args):
    import torch.distributed as dist

    from glass_b200 import build as _build
    _build.build()
    from glass_b200 import ops, run
    from glass_b200.graphed import GraphedTrainStep, train_epoch
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = make_workload(args.workload)
    p, g = wl["params"], wl["g"]
    torch.manual_seed(0)
    model = run.build_model(p["hidden_dim"], p["conv_layer"], p["dropout"], 1, p["pool"], p["z_ratio"], p["aggr"],
                            wl["max_deg"], wl["out_dim"], pretrained=wl["table"], device=dev)
    x, ei, ew = g.x.to(dev), g.edge_index.to(dev), g.edge_attr.to(dev)
    loss_fn = wl["loss_fn"]
    n_total = args.warmup + args.steps
    host_batches = [(pos.pin_memory(), y.pin_memory()) for pos, y in batches_for(wl, n_total, rank, world)]
    dev_batches = [(pos.to(dev), y.to(dev)) for pos, y in host_batches]
    # one captured CUDA graph per step: labels, forward, loss, backward, (NCCL all-reduce), Adam
    ops.reset_launch_count()
    step = GraphedTrainStep(model, loss_fn, x, ei, ew, dev_batches[0][0], dev_batches[0][1], p["lr"], warmup=3)
    ops.reset_launch_count()
    if not args.no_graph:
        step.capture()
    else:
        step(*dev_batches[0])
    launches_per_step = ops.launch_count()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-timed value: inputs resident in HBM
    model.train()
    for pos, y in dev_batches[:args.warmup]:
        step(pos, y)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        e0.record()
        for pos, y in dev_batches[args.warmup:]:
            step(pos, y)
        e1.record()
        barrier()
    launches = launches_per_step * args.steps
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    bs = p["batch_size"]
    value = bs * args.steps * world / (ms * 1e-3)

    # ---- e2e: public train API, host buffers -> device every step, loss read back every step
    barrier()
    t0 = time.perf_counter()
    train_epoch(step, host_batches[args.warmup:], sync_each_step=True)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = bs * args.steps * world / float(e2e_s)
    h2d = host_batches[0][0].numel() * 8 + host_batches[0][1].numel() * host_batches[0][1].element_size()

    if rank == 0:
        line = base_line(args, wl, value, ms / args.steps)
        line["e2e"] = {"value": e2e_value, "unit": "subgraphs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                       "api": "glass_b200.graphed.train_epoch(GraphedTrainStep, pinned host batches), loss.item() every step"}
        line["cuda_graph"] = not args.no_graph
        if world > 1:   # "symm": fused reduce-scatter + Adam + all-gather over NVLink peer memory; "nccl": all-reduce, then Adam
            line["grad_exchange"] = step.dp_mode
        line["gpu_launches"] = launches
        line["clocks"] = clocks.summary()
        adj = model.conv.convs[0].adj
        line["roofline"] = spmm_roofline(adj, p["hidden_dim"], workload=wl["name"])
        # inference throughput of the same model / batches (impl/train.py:20-34 forward only): one graph per batch,
        # and the multi-label-batch evaluator (adj @ U once per epoch + sparse label correction per batch)
        from glass_b200.graphed import GraphedForward, GraphedSharedBaseForward
        fwd = GraphedForward(model, x, ei, ew, dev_batches[0][0])
        for pos, y in dev_batches[:args.warmup]:
            fwd(pos)
        torch.cuda.synchronize()
        e0.record()
        for pos, y in dev_batches[args.warmup:]:
            fwd(pos)
        e1.record()
        torch.cuda.synchronize()
        line["infer"] = {"value": bs * args.steps / (e0.elapsed_time(e1) * 1e-3), "unit": "subgraphs/s",
                         "n_gpus": 1, "ms_per_step": e0.elapsed_time(e1) / args.steps}
        try:
            sb = GraphedSharedBaseForward(model, x, ei, ew, dev_batches[0][0])
            per_epoch = max(1, int(wl["g"].pos.shape[0] - wl["trn_pos"].shape[0]) // bs)   # val + test batches
            for pos, y in dev_batches[:args.warmup]:
                sb(pos)
            torch.cuda.synchronize()
            e0.record()
            for i, (pos, y) in enumerate(dev_batches[args.warmup:]):
                if i % per_epoch == 0:
                    sb.refresh()                    # new weights every evaluation epoch
                sb(pos)
            e1.record()
            torch.cuda.synchronize()
            ms_sb = e0.elapsed_time(e1)
            e0.record()
            for pos, y in dev_batches[args.warmup:]:
                sb(pos)
            e1.record()
            torch.cuda.synchronize()
            line["infer_shared_base"] = {
                "value": bs * args.steps / (ms_sb * 1e-3), "unit": "subgraphs/s", "ms_per_step": ms_sb / args.steps,
                "refresh_every_batches": per_epoch, "ms_per_step_steady": e0.elapsed_time(e1) / args.steps,
                "what": "glass_b200.graphed.GraphedSharedBaseForward: label-independent base once per evaluation epoch "
                        "(val + test split), sparse label correction per batch; same logits as `infer`"}
            del sb
        except NotImplementedError as e:
            line["infer_shared_base"] = {"unavailable": str(e)[:200]}
        if world == 1 and not args.no_kernel_rooflines:
            try:
                line["kernel_rooflines"] = kernel_rooflines(model, wl, x, ei, ew, dev_batches[0][0], dev,
                                                            line["roofline"]["peak"])
            except Exception as e:  # noqa: BLE001  (secondary numbers must not take the headline line down)
                line["kernel_rooflines"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        if world == 1 and not args.no_cpu_baseline:
            v, per = cpu_port_steps(wl, args.cpu_steps, 1)
            line["cpu_baseline"] = {"value": v, "unit": "subgraphs/s", "cores": torch.get_num_threads(),
                                    "kind": "port",
                                    "sample": f"{args.cpu_steps} train steps (after 1 warm-up) of batch {bs} on the full "
                                              f"{wl['name']} graph, oracle port of impl/models.py",
                                    "ms_per_step": per * 1e3}
        if world == 1 and not args.no_gpu_eager_baseline:
            v, per = cpu_port_steps(wl, 10, 2, device=dev)
            line["gpu_eager_baseline"] = {"value": v, "unit": "subgraphs/s", "ms_per_step": per * 1e3,
                                          "what": "oracle port of impl/models.py run eagerly on this GPU "
                                                  "(torch.sparse COO @ dense -> cuSPARSE, ATen element-wise, torch Adam)"}
        if world == 1 and args.workload == "em_user_shaped" and not args.no_other_configs:
            del fwd, step
            line["other_configs"] = {}
            for name in OTHER_CONFIGS:      # secondary numbers: a failure here must not take the headline line down
                try:
                    line["other_configs"][name] = quick_config(name, dev, args.steps, args.warmup,
                                                               0 if args.no_cpu_baseline else args.cpu_steps)
                except Exception as e:  # noqa: BLE001
                    line["other_configs"][name] = {"error": f"{type(e).__name__}: {e}"[:300]}
        print(json.dumps(line), flush=True)
    if world > 1:
        # CUDA graphs that captured NCCL kernels are still alive; tearing the communicator down under them
        # can hang, so every rank synchronises its device, flushes and leaves without running destructors
        # (no collective is pending: the last one was the e2e max-reduce above).
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


# BASELINE.json configs[0..2] + the skewed variant of configs[3] ("use a skewed /
# power-law degree generator as well as uniform")
OTHER_CONFIGS = ("density", "cut_ratio", "component", "ppi_bp_shaped", "em_user_shaped_powerlaw")


def quick_config(name, dev, steps, warmup, cpu_steps):
    """Device-timed graph-replayed train steps of another BASELINE.json config (batches resident in HBM) and
    the CPU port beside it; reported under `other_configs`, not part of the headline value."""
    from glass_b200 import ops, run
    from glass_b200.graphed import GraphedTrainStep
    wl = make_workload(name)
    p, g = wl["params"], wl["g"]
    torch.manual_seed(0)
    model = run.build_model(p["hidden_dim"], p["conv_layer"], p["dropout"], 1, p["pool"], p["z_ratio"], p["aggr"],
                            wl["max_deg"], wl["out_dim"], pretrained=wl["table"], device=dev)
    x, ei, ew = g.x.to(dev), g.edge_index.to(dev), g.edge_attr.to(dev)
    batches = [(pos.to(dev), y.to(dev)) for pos, y in batches_for(wl, warmup + steps, 0, 1)]
    ops.reset_launch_count()
    step = GraphedTrainStep(model, wl["loss_fn"], x, ei, ew, batches[0][0], batches[0][1], p["lr"], warmup=3)
    ops.reset_launch_count()
    step.capture()
    launches = ops.launch_count()
    for pos, y in batches[:warmup]:
        step(pos, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for pos, y in batches[warmup:]:
        step(pos, y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out = {"value": p["batch_size"] / (ms * 1e-3), "unit": "subgraphs/s", "ms_per_step": ms,
           "gpu_launches_per_step": launches, "nodes": int(g.x.shape[0]), "nnz": int(model.conv.convs[0].adj.nnz),
           "hidden_dim": p["hidden_dim"], "conv_layer": p["conv_layer"], "batch_size": p["batch_size"]}
    if name.startswith("em_user"):
        out["roofline"] = spmm_roofline(model.conv.convs[0].adj, p["hidden_dim"], workload=name)
    if cpu_steps and not name.startswith("em_user"):
        v, per = cpu_port_steps(wl, cpu_steps, 1)
        out["cpu_port"] = {"value": v, "ms_per_step": per * 1e3, "cores": torch.get_num_threads()}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="glass_b200", choices=["glass_b200", "reference"])
    ap.add_argument("--workload", default="em_user_shaped")
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true",
                    help="skip the short runs of the other BASELINE.json configs (`other_configs` key)")
    ap.add_argument("--no-kernel-rooflines", action="store_true",
                    help="skip the isolated timings of csr_build / GraphNorm / pair GEMMs / pooling / Adam (`kernel_rooflines`)")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of one CUDA graph per step")
    ap.add_argument("--no-gpu-eager-baseline", action="store_true",
                    help="skip timing the reference's eager torch.sparse op sequence on this GPU (`gpu_eager_baseline` key)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl != "reference":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
