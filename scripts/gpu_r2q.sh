#!/bin/bash
set -x
O=gpurun_out/r2q
mkdir -p $O
python -c "import glass_b200.build as b; print(b.build())" > $O/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout=600 -k "spmm" > $O/pytest_spmm.log 2>&1
HUB_OUT=r2q/hub_check.json timeout 900 python scripts/hub_check.py em_user_shaped_powerlaw em_user_shaped stress > $O/hub_check.log 2>&1
GLASS_B200_TC_TIMELINE=1 timeout 300 python scripts/tc_timeline.py > $O/tc_timeline.log 2>&1
tail -5 $O/pytest_spmm.log; cat $O/hub_check.log
