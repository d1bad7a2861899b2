#!/bin/bash
# label-batch DP scaling on one 8-GPU box with the final build: em_user-shaped at N = 1, 2, 4, 8; ppi_bp-shaped at N = 1, 8
set -x
O=gpurun_out/$1
mkdir -p $O
python -c "import glass_b200.build as b; print(b.build())" > $O/build.log 2>&1
FAST="--no-other-configs --no-cpu-baseline --no-kernel-rooflines --no-gpu-eager-baseline"
timeout 300 python bench.py --gpus 1 --steps 100 --warmup 10 $FAST > $O/bench_em_user_n1.json 2> $O/bench_em_user_n1.err
for N in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 100 --warmup 10 $FAST > $O/bench_em_user_n$N.json 2> $O/bench_em_user_n$N.err
done
timeout 300 python bench.py --gpus 1 --steps 100 --warmup 10 --workload ppi_bp_shaped $FAST > $O/bench_ppi_bp_n1.json 2> $O/bench_ppi_bp_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 100 --warmup 10 --workload ppi_bp_shaped $FAST > $O/bench_ppi_bp_n8.json 2> $O/bench_ppi_bp_n8.err
for f in $O/bench_*.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['n_gpus'], round(d['value']), d['ms_per_step'], d.get('grad_exchange'))"; done
tail -n 2 $O/*.err
