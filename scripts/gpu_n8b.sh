#!/bin/bash
set -x
O=gpurun_out/$1
mkdir -p $O
python -c "import glass_b200.build as b; print(b.build())" > $O/build.log 2>&1
timeout 300 python scripts/profile_step.py > $O/warm_em_user.txt 2>&1
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR8 scripts/bench_stress.py --graph stress --pipelined --check > $O/stress_pipelined_n8.json 2> $O/stress_pipelined_n8.err
GLASS_B200_PARTITION_PHASES=3 timeout 300 $TR8 scripts/bench_stress.py --graph stress --pipelined > $O/stress_pipelined3_n8.json 2> $O/stress_pipelined3_n8.err
GLASS_B200_PARTITION_PHASES=2 timeout 300 $TR8 scripts/bench_stress.py --graph stress --pipelined > $O/stress_pipelined2_n8.json 2> $O/stress_pipelined2_n8.err
tail -c 500 $O/*.json; tail -n 3 $O/*.err; head -20 $O/warm_em_user.txt
