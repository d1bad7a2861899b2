#!/bin/bash
set -x
O=gpurun_out/r2e
mkdir -p $O
python -c "import glass_b200.build as b; print(b.build())" > $O/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -x > $O/pytest.log 2>&1
echo "pytest rc $?" >> $O/pytest.log
timeout 300 python scripts/gemm_time.py > $O/gemm_time.txt 2>&1
timeout 600 python bench.py --no-other-configs --no-cpu-baseline --steps 50 > $O/bench.json 2> $O/bench.err
timeout 300 python scripts/profile_step.py > $O/warm_em_user.txt 2>&1
# kernel order per shape: fwd, dX, dW, reduce ; skip the first round (9 launches incl. reduce)


ls -la $O
