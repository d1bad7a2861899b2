#!/bin/bash
# usage: gpu_multi.sh N TAG   -- label-batch DP bench (configs 3, 4) and the row-partitioned stress runs (config 5) on N GPUs
set -x
N=$1; O=gpurun_out/$2
mkdir -p $O
python -c "import glass_b200.build as b; print(b.build())" > $O/build.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR bench.py --gpus $N --steps 100 --warmup 10 > $O/bench_em_user_n$N.json 2> $O/bench_em_user_n$N.err
timeout 400 $TR bench.py --gpus $N --steps 100 --warmup 10 --workload ppi_bp_shaped > $O/bench_ppi_bp_n$N.json 2> $O/bench_ppi_bp_n$N.err
timeout 400 $TR scripts/bench_stress.py --graph stress --check > $O/stress_n$N.json 2> $O/stress_n$N.err
timeout 400 $TR scripts/bench_stress.py --graph stress --overlap > $O/stress_overlap_n$N.json 2> $O/stress_overlap_n$N.err
timeout 300 $TR scripts/partition_check.py stress_small > $O/partition_check_n$N.json 2> $O/partition_check_n$N.err
tail -c 400 $O/*.json; tail -n 3 $O/*.err
