#!/bin/bash
set -x
O=gpurun_out/$1
mkdir -p $O
python -c "import glass_b200.build as b; print(b.build())" > $O/build.log 2>&1
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -q -m gpu --timeout=1400 > $O/memcheck.log 2>&1; echo "memcheck rc $?" >> $O/memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout=1400 -k "graph_norm or spmm or segment_pool or pair_linear or norm_pool or embedding or row_partitioned or to_undirected" > $O/racecheck.log 2>&1; echo "racecheck rc $?" >> $O/racecheck.log
tail -5 $O/memcheck.log $O/racecheck.log
