#!/bin/bash
# last call of the round on the committed tree: full GPU suite, smoke, default bench, reference arm
set -x
O=gpurun_out/$1
mkdir -p $O
python -c "import glass_b200.build as b; print(b.build())" > $O/build.log 2>&1
timeout 1800 python -m pytest tests -m gpu -x -q --timeout=900 > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc $?" >> $O/smoke.log
timeout 1500 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc $?" >> $O/bench_default.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
for w in ppi_bp_shaped coreness; do timeout 300 python scripts/profile_step.py $w > $O/step_warm_kernel_times_$w.txt 2>&1; done
tail -n 3 $O/pytest.log $O/smoke.log; tail -n 1 $O/bench_default.err
