#!/bin/bash
set -x
O=gpurun_out/r2s
mkdir -p $O
python -c "import glass_b200.build as b; print(b.build())" > $O/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_fullsize.py -m gpu -q --timeout=600 -k "pair or model or full or conv or golden" > $O/pytest.log 2>&1
for e in 8 16; do
GLASS_B200_TC_EPI=$e timeout 300 python scripts/gemm_time.py > $O/gemm_time_epi$e.txt 2>&1
done
GLASS_B200_TC_TIMELINE=1 timeout 300 python scripts/tc_timeline.py > $O/tc_timeline.log 2>&1
timeout 300 python scripts/profile_step.py > $O/warm_em_user.txt 2>&1
tail -4 $O/pytest.log; cat $O/gemm_time_epi8.txt $O/gemm_time_epi16.txt; head -12 $O/tc_timeline.log; head -20 $O/warm_em_user.txt
