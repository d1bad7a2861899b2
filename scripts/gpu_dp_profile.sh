#!/bin/bash
set -x
O=gpurun_out/$1
mkdir -p $O
python -c "import glass_b200.build as b; print(b.build())" > $O/build.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR scripts/profile_step_dp.py > $O/dp_profile_symm_n2.txt 2> $O/dp_profile_symm_n2.err
GLASS_B200_DP=nccl timeout 300 $TR scripts/profile_step_dp.py > $O/dp_profile_nccl_n2.txt 2> $O/dp_profile_nccl_n2.err
timeout 300 python scripts/profile_step_dp.py > $O/dp_profile_n1.txt 2> $O/dp_profile_n1.err
head -30 $O/dp_profile_symm_n2.txt; head -8 $O/dp_profile_nccl_n2.txt; head -3 $O/dp_profile_n1.txt; tail -3 $O/*.err
