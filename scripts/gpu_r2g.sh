#!/bin/bash
set -x
O=gpurun_out/r2g
mkdir -p $O
python -c "import glass_b200.build as b; print(b.build())" > $O/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -x > $O/pytest.log 2>&1
echo "pytest rc $?" >> $O/pytest.log
GLASS_B200_TC_TIMELINE=1 timeout 300 python scripts/tc_timeline.py > $O/tc_timeline.log 2>&1
timeout 300 python scripts/gemm_time.py > $O/gemm_time.txt 2>&1
for fused in 1 0; do for coop in 1 0; do
  GLASS_B200_CONV_FUSED=$fused GLASS_B200_GN_COOP=$coop timeout 300 python scripts/profile_step.py > $O/warm_fused${fused}_coop${coop}.txt 2>&1
done; done
GLASS_B200_CONV_FUSED=0 timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -x > $O/pytest_unfused.log 2>&1
echo "pytest rc $?" >> $O/pytest_unfused.log
timeout 600 python bench.py --no-other-configs --no-cpu-baseline --no-gpu-eager-baseline --steps 50 > $O/bench.json 2> $O/bench.err
ls -la $O
