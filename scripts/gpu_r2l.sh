#!/bin/bash
set -x
O=gpurun_out/r2l
mkdir -p $O
python -c "import glass_b200.build as b; print(b.build())" > $O/build.log 2>&1
timeout 300 python scripts/debug_edgegnn.py > $O/debug_edgegnn.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout=900 -k "pretraining or reproducible" > $O/pytest_sub.log 2>&1
ls -la $O
