#!/bin/bash
set -x
O=gpurun_out/r2k
mkdir -p $O
python -c "import glass_b200.build as b; print(b.build())" > $O/build.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q --timeout=900 > $O/pytest.log 2>&1
echo "pytest rc $?" >> $O/pytest.log
timeout 300 python scripts/profile_step.py > $O/warm_em_user.txt 2>&1
timeout 300 python scripts/profile_step.py density > $O/warm_density.txt 2>&1
timeout 300 python scripts/profile_step.py ppi_bp_shaped > $O/warm_ppi.txt 2>&1
ls -la $O
