#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (+ optional ncu passes).  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
if [ -n "$TC_DEBUG" ]; then
  timeout -s KILL 240 python scripts/tc_debug.py > gpurun_out/tc_debug.log 2>&1; echo "tc_debug exit $?" >> gpurun_out/tc_debug.log
  cat gpurun_out/tc_debug.log | tail -40
fi
GLASS_B200_GEMM=${GEMM:-auto} timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 600 ${PYTEST_ARGS} > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -${PYTEST_TAIL:-25} gpurun_out/pytest_gpu.log
GLASS_B200_GEMM=${GEMM:-auto} timeout -s KILL 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
if [ -z "$SKIP_BENCH" ]; then
  GLASS_B200_GEMM=${GEMM:-auto} timeout -s KILL 900 python bench.py ${BENCH_ARGS} > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
  tail -3 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
fi
