#!/bin/bash
# round-2 GPU call A: full GPU suite (incl. the new full-size parity tests), default bench, ncu evidence for k_spmm
set -x
O=gpurun_out/r2a
mkdir -p $O
nvidia-smi > $O/gpu.txt 2>&1
python -c "import glass_b200.build as b; print(b.build())" > $O/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 > $O/pytest.log 2>&1
echo "pytest rc $?" >> $O/pytest.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
timeout 300 python scripts/profile_step.py > $O/warm_em_user.txt 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_spmm -s 2 -c 1 -f -o $O/spmm_uniform python scripts/spmm_time.py em_user_shaped > $O/ncu_spmm_uniform.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_spmm -s 4 -c 2 -f -o $O/spmm_powerlaw python scripts/spmm_time.py em_user_shaped_powerlaw > $O/ncu_spmm_powerlaw.log 2>&1
timeout 300 python bench.py --workload em_user_shaped_powerlaw --no-other-configs --no-cpu-baseline --steps 50 > $O/bench_powerlaw.json 2> $O/bench_powerlaw.err
ls -la $O
