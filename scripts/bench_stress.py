#!/usr/bin/env python
"""Row-partitioned SpMM on the stress graph (BASELINE.json configs[4]): power-law, 2M nodes, 100M undirected
edges (nnz 200M), H = 64; one process per GPU under torchrun, NCCL all-gather of the feature shards + local
block SpMM.  Prints one JSON line per rank-0 run:  python -m torch.distributed.run --nproc-per-node P \
scripts/bench_stress.py [--graph stress|stress_small] [--check]"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--graph", default="stress_small")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--check", action="store_true", help="compare with the replicated single-GPU SpMM")
    ap.add_argument("--overlap", action="store_true", help="multiply locally-owned columns while the all-gather runs")
    ap.add_argument("--pipelined", action="store_true",
                    help="peer-memory pulls per source rank, one accumulating SpMM phase per arrived shard")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from glass_b200 import datasets, ops
    from glass_b200.partition import RowPartitionedAdj
    t0 = time.time()
    g = datasets.load_dataset(args.graph, device=dev)      # generated on the GPU (same seed on every rank)
    n, h = g.num_nodes, 64
    t_gen = time.time() - t0
    t0 = time.time()
    adj = ops.build_csr(g.edge_index.to(dev), g.edge_attr.to(dev), n, "gcn")
    torch.cuda.synchronize()
    t_build = time.time() - t0
    part = RowPartitionedAdj(adj, rank, world, overlap=args.overlap, pipelined=args.pipelined)
    x = torch.randn(n, h, device=dev, generator=torch.Generator(device=dev).manual_seed(0))
    xs = part.shard(x)
    err = None
    if args.check:
        y_ref = ops.spmm(adj, x)[part.lo:part.hi]
        y = part.spmm(xs)
        err = float((y - y_ref).abs().max() / y_ref.abs().max())
    if not args.check or True:
        del adj      # keep only the block
    torch.cuda.empty_cache()
    for _ in range(3):
        part.spmm(xs)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    gather_ms = spmm_ms = 0.0
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        part.spmm(xs)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    # split: gather alone
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(args.steps):
        part._gather(xs)
    g1.record()
    torch.cuda.synchronize()
    gms = torch.tensor([g0.elapsed_time(g1) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(gms, op=dist.ReduceOp.MAX)
    if rank == 0:
        nnz = int(g.edge_index.shape[1])
        algo = 4 * (n + 1) + 8 * nnz + 8 * n * h
        print(json.dumps({"metric": "row-partitioned SpMM (all-gather + local block SpMM)", "graph": args.graph,
                          "overlap": args.overlap, "pipelined": args.pipelined, "nodes": n, "nnz": nnz, "h": h, "n_gpus": world, "ms_per_spmm": float(ms),
                          "ms_allgather": float(gms), "algorithmic_GBps_aggregate": algo / float(ms) / 1e6,
                          "rows_per_rank_pad": part.pad, "nnz_rank0": part.nnz_local, "check_rel_err": err,
                          "gen_s": t_gen, "csr_build_s": t_build}), flush=True)
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
