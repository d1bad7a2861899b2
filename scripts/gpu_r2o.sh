#!/bin/bash
# N=2 sanity of every multi-GPU entry point before spending 4x / 8x minutes
set -x
O=gpurun_out/r2o
mkdir -p $O
python -c "import glass_b200.build as b; print(b.build())" > $O/build.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR scripts/partition_check.py stress_small > $O/partition_check_n2.json 2> $O/partition_check_n2.err
timeout 300 $TR scripts/bench_stress.py --graph stress_small --check > $O/stress_small_n2.json 2> $O/stress_small_n2.err
timeout 400 $TR bench.py --gpus 2 --steps 50 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err
timeout 400 $TR bench.py --gpus 2 --steps 50 --warmup 5 --workload ppi_bp_shaped > $O/bench_ppi_n2.json 2> $O/bench_ppi_n2.err
timeout 600 python -m pytest tests -m gpu -q --timeout=600 -k "dist or nccl or world or partition" > $O/pytest_multi.log 2>&1
tail -c 600 $O/*.json; tail -n 5 $O/*.err $O/pytest_multi.log
