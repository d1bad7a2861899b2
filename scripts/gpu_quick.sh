#!/bin/bash
# quick single-GPU check of a kernel change: targeted tests ($2 = pytest -k expression), warm step profile, fast bench
set -x
O=gpurun_out/$1
mkdir -p $O
python -c "import glass_b200.build as b; print(b.build())" > $O/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout=600 -k "$2" > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log
timeout 300 python scripts/profile_step.py > $O/warm_em_user.txt 2>&1
timeout 300 python scripts/profile_step.py ppi_bp_shaped > $O/warm_ppi.txt 2>&1
timeout 300 python bench.py --no-other-configs --no-cpu-baseline --no-kernel-rooflines --no-gpu-eager-baseline > $O/bench_fast.json 2> $O/bench_fast.err
tail -n 3 $O/pytest.log; head -n 16 $O/warm_em_user.txt | tail -n 14; sed -n 3p $O/warm_ppi.txt
python -c "
import json
d=json.loads(open('$O/bench_fast.json').read().strip().splitlines()[-1])
print('train ms', round(d['ms_per_step'],4), 'infer', round(d['infer']['ms_per_step'],4))"
