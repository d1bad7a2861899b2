#!/bin/bash
set -x
O=gpurun_out/r2m
mkdir -p $O
python -c "import glass_b200.build as b; print(b.build())" > $O/build.log 2>&1
T=tests/test_gpu_model.py::test_pretraining_modules_match_reference_golden
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout=900 2>&1 | tail -5 > $O/a_model_only.log
timeout 900 python -m pytest tests/test_gpu_kernels.py $T -m gpu -q --timeout=900 2>&1 | tail -5 > $O/b_kernels_then.log
timeout 900 python -m pytest tests/test_gpu_fullsize.py $T -m gpu -q --timeout=900 2>&1 | tail -5 > $O/c_fullsize_then.log
timeout 900 python -m pytest tests/test_gpu_partition.py $T -m gpu -q --timeout=900 2>&1 | tail -5 > $O/d_partition_then.log
tail -n 3 $O/*.log
