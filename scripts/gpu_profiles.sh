#!/bin/bash
# full GPU suite, smoke, warm step profiles, ncu evidence (launch list + --set full captures), 30-seed protocol runs
set -x
O=gpurun_out/$1
mkdir -p $O
nvidia-smi > $O/gpu.txt 2>&1
python -c "import glass_b200.build as b; print(b.build())" > $O/build.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q --timeout=900 > $O/pytest.log 2>&1
echo "pytest rc $?" >> $O/pytest.log
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc $?" >> $O/smoke.log
for w in em_user_shaped em_user_shaped_powerlaw ppi_bp_shaped density cut_ratio component coreness; do
  timeout 300 python scripts/profile_step.py $w > $O/step_warm_kernel_times_$w.txt 2>&1
done
timeout 600 python scripts/spmm_probe.py em_user_shaped em_user_shaped_powerlaw stress > $O/spmm_probe.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 400 $NCU -k regex:k_spmm -s 2 -c 1 -o $O/spmm_uniform python scripts/spmm_time.py em_user_shaped > $O/ncu_spmm_uniform.log 2>&1
timeout 400 $NCU -k regex:"k_spmm" -s 4 -c 2 -o $O/spmm_powerlaw python scripts/spmm_time.py em_user_shaped_powerlaw > $O/ncu_spmm_powerlaw.log 2>&1
timeout 400 $NCU -k regex:"k_gn_coop" -s 4 -c 2 -o $O/gn_coop python scripts/gn_ncu.py > $O/ncu_gn_coop.log 2>&1
timeout 400 $NCU -k regex:"k_pool_pad|k_colsums|k_gn_finalize|k_mark_nodes" -s 10 -c 5 -o $O/norm_pool python scripts/gn_pool_ncu.py > $O/ncu_norm_pool.log 2>&1
timeout 400 $NCU -k regex:"k_pair_tc|k_pair_dw_tc" -s 16 -c 8 -o $O/pair_gemm python scripts/tc_ncu.py > $O/ncu_pair_gemm.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/step_launches_em_user_graph.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs --no-kernel-rooflines --no-gpu-eager-baseline > $O/bench_under_ncu.log 2>&1
for ds in density cut_ratio component coreness; do
  timeout 900 python GLASSTest.py --use_one --use_seed --use_maxzeroone --repeat 30 --device 0 --dataset $ds --graph > $O/glasstest_repeat30_$ds.log 2>&1
done
ls -la $O; tail -3 $O/pytest.log $O/smoke.log
