#!/bin/bash
set -x
O=gpurun_out/$1
mkdir -p $O
python -c "import glass_b200.build as b; print(b.build())" > $O/build.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518"
timeout 300 $TR scripts/bench_stress.py --graph stress --pipelined --check > $O/stress_pipelined_n8.json 2> $O/stress_pipelined_n8.err
timeout 300 $TR scripts/bench_stress.py --graph stress > $O/stress_n8.json 2> $O/stress_n8.err
for f in $O/stress_*.json; do grep -v NCCL $f | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$f', d['n_gpus'], d['pipelined'], round(d['ms_per_spmm'],3), round(d['ms_allgather'],3), d['check_rel_err'])"; done
