"""torchrun --nproc-per-node P scripts/allreduce_bench.py : NCCL vs symmetric-memory all-reduce of the DP gradient buffer."""
import os, sys
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as sm
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 3_700_000  # ~ number of fp32 gradients at the em_user shape
n = (n + 1023) // 1024 * 1024
x = torch.randn(n, device=dev)
buf = sm.empty(n, dtype=torch.float32, device=dev)
hdl = sm.rendezvous(buf, dist.group.WORLD)
gname = dist.group.WORLD.group_name
if rank == 0:
    print("multicast_ptr", hdl.multicast_ptr, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs])

def timeit(fn, reps=30, graph=False):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    if graph:
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s): fn()
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(g): fn()
        run = g.replay
    else:
        run = fn
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): run()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps * 1e3], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)

tests = {
    "nccl all_reduce": lambda: dist.all_reduce(x),
    "symm two_shot": lambda: torch.ops.symm_mem.two_shot_all_reduce_(buf, "sum", gname),
    "symm one_shot": lambda: torch.ops.symm_mem.one_shot_all_reduce(buf, "sum", gname),
}
if hdl.multicast_ptr:
    tests["symm multimem"] = lambda: torch.ops.symm_mem.multimem_all_reduce_(buf, "sum", gname)
for name, fn in tests.items():
    for graph in (False, True):
        try:
            buf.copy_(x)
            us = timeit(fn, graph=graph)
            if rank == 0: print(f"{name:18s} graph={graph}: {us:8.1f} us for {n*4/1e6:.1f} MB, world {world}", flush=True)
        except Exception as e:
            if rank == 0: print(name, "graph", graph, "ERR", type(e).__name__, str(e)[:200], flush=True)
# correctness of two_shot vs nccl
buf.copy_(x); torch.ops.symm_mem.two_shot_all_reduce_(buf, "sum", gname)
y = x.clone(); dist.all_reduce(y)
if rank == 0: print("two_shot max abs diff vs nccl:", float((buf - y).abs().max()))
torch.cuda.synchronize(); sys.stdout.flush(); os._exit(0)
