#!/bin/bash
# final call B: GPU suite on the final kernels, warm step profiles, ncu --set full of k_spmm (traffic stamp), combine / pooled-backward captures
set -x
O=gpurun_out/$1
mkdir -p $O
python -c "import glass_b200.build as b; print(b.build())" > $O/build.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q --timeout=900 > $O/pytest.log 2>&1
echo "pytest rc $?" >> $O/pytest.log
for w in em_user_shaped em_user_shaped_powerlaw ppi_bp_shaped density cut_ratio component coreness; do
  timeout 300 python scripts/profile_step.py $w > $O/step_warm_kernel_times_$w.txt 2>&1
done
timeout 600 python scripts/spmm_probe.py em_user_shaped em_user_shaped_powerlaw stress > $O/spmm_probe.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 400 $NCU -k regex:k_spmm -s 2 -c 1 -o $O/spmm_uniform python scripts/spmm_time.py em_user_shaped > $O/ncu_spmm_uniform.log 2>&1
timeout 400 $NCU -k regex:"k_spmm" -s 4 -c 2 -o $O/spmm_powerlaw python scripts/spmm_time.py em_user_shaped_powerlaw > $O/ncu_spmm_powerlaw.log 2>&1
timeout 400 $NCU -k regex:"k_pool_pad|k_colsums|k_gn_finalize|k_mark_nodes" -s 10 -c 5 -o $O/norm_pool python scripts/gn_pool_ncu.py > $O/ncu_norm_pool.log 2>&1
tail -3 $O/pytest.log; head -20 $O/step_warm_kernel_times_em_user_shaped.txt; cat $O/spmm_probe.log
