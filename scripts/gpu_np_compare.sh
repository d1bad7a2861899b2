#!/bin/bash
set -x
O=gpurun_out/$1
mkdir -p $O
python -c "import glass_b200.build as b; print(b.build())" > $O/build.log 2>&1
FAST="--no-other-configs --no-cpu-baseline --no-kernel-rooflines --no-gpu-eager-baseline"
for v in 1 0 1 0; do
GLASS_B200_NORM_POOL=$v timeout 300 python bench.py $FAST > $O/bench_np$v.json 2> $O/bench_np$v.err
python -c "
import json
d=json.loads(open('$O/bench_np$v.json').read().strip().splitlines()[-1])
print('NORM_POOL=$v train', round(d['ms_per_step'],4), 'infer', round(d['infer']['ms_per_step'],4), 'shared', round(d['infer_shared_base']['ms_per_step'],4), round(d['infer_shared_base']['ms_per_step_steady'],4))"
done
