#!/bin/bash
# 8-GPU box: label-batch DP (configs 3, 4) at N=8, row-partitioned stress graph (config 5) at N=8 and N=4 (pipelined / plain),
# the row-partitioned model check at N=8
set -x
O=gpurun_out/$1
mkdir -p $O
python -c "import glass_b200.build as b; print(b.build())" > $O/build.log 2>&1
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512"
timeout 300 $TR8 scripts/bench_stress.py --graph stress --pipelined --check > $O/stress_pipelined_n8.json 2> $O/stress_pipelined_n8.err
timeout 300 $TR8 scripts/bench_stress.py --graph stress > $O/stress_n8.json 2> $O/stress_n8.err
timeout 300 $TR4 scripts/bench_stress.py --graph stress --pipelined --check > $O/stress_pipelined_n4.json 2> $O/stress_pipelined_n4.err
timeout 400 $TR8 bench.py --gpus 8 --steps 100 --warmup 10 > $O/bench_em_user_n8.json 2> $O/bench_em_user_n8.err
timeout 400 $TR8 bench.py --gpus 8 --steps 100 --warmup 10 --workload ppi_bp_shaped > $O/bench_ppi_bp_n8.json 2> $O/bench_ppi_bp_n8.err
PIPELINED=1 timeout 300 $TR8 scripts/partition_check.py stress_small > $O/partition_check_pipelined_n8.json 2> $O/partition_check_pipelined_n8.err
tail -c 500 $O/*.json; tail -n 3 $O/*.err
