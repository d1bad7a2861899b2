#!/bin/bash
set -x
O=gpurun_out/$1
mkdir -p $O
python -c "import glass_b200.build as b; print(b.build())" > $O/build.log 2>&1
FAST="--no-other-configs --no-cpu-baseline --no-kernel-rooflines --no-gpu-eager-baseline"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus 2 --steps 100 --warmup 10 $FAST > $O/bench_em_user_n2.json 2> $O/bench_em_user_n2.err
timeout 300 $TR bench.py --gpus 2 --steps 100 --warmup 10 --workload ppi_bp_shaped $FAST > $O/bench_ppi_bp_n2.json 2> $O/bench_ppi_bp_n2.err
timeout 300 $TR scripts/dp_check.py > $O/dp_check.json 2> $O/dp_check.err
for f in $O/bench_*.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['n_gpus'], round(d['value']), round(d['ms_per_step'],4), d.get('grad_exchange'))"; done
tail -c 600 $O/dp_check.json
