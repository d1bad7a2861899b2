#!/bin/bash
set -x
O=gpurun_out/r2r
mkdir -p $O
python -c "import glass_b200.build as b; print(b.build())" > $O/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_partition.py -m gpu -q --timeout=600 -k "spmm or partition" > $O/pytest.log 2>&1
timeout 600 python scripts/spmm_probe.py em_user_shaped em_user_shaped_powerlaw > $O/spmm_probe.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR scripts/bench_stress.py --graph stress_small --pipelined --check > $O/stress_small_pipe_n2.json 2> $O/stress_small_pipe_n2.err
timeout 400 $TR scripts/bench_stress.py --graph stress --pipelined --check > $O/stress_pipe_n2.json 2> $O/stress_pipe_n2.err
timeout 400 $TR scripts/bench_stress.py --graph stress > $O/stress_n2.json 2> $O/stress_n2.err
tail -5 $O/pytest.log; cat $O/spmm_probe.log; tail -c 700 $O/*.json; tail -n 4 $O/*.err
