#!/bin/bash
set -x
O=gpurun_out/r2u
mkdir -p $O
python -c "import glass_b200.build as b; print(b.build())" > $O/build.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q --timeout=900 > $O/pytest.log 2>&1
echo "pytest rc $?" >> $O/pytest.log
timeout 300 python scripts/profile_step.py > $O/warm_em_user.txt 2>&1
GLASS_B200_NORM_POOL=0 timeout 300 python scripts/profile_step.py > $O/warm_em_user_nofuse.txt 2>&1
timeout 300 python scripts/profile_step.py ppi_bp_shaped > $O/warm_ppi.txt 2>&1
timeout 300 python scripts/profile_step.py density > $O/warm_density.txt 2>&1
timeout 600 python scripts/spmm_probe.py em_user_shaped_powerlaw > $O/spmm_probe.log 2>&1
tail -8 $O/pytest.log; head -24 $O/warm_em_user.txt; head -8 $O/warm_em_user_nofuse.txt; cat $O/spmm_probe.log
