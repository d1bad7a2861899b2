#!/bin/bash
# DP step time at N GPUs under a few NCCL settings (the all-reduce of the flat gradient is captured in the step graph)
N=${1:-2}
run() {
  env "$@" timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $N --steps 50 --warmup 5 2>/dev/null | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value'],1))
except Exception as e: print('$*', 'ERR', e)"
}
run X=default
run NCCL_MIN_NCHANNELS=16
run NCCL_MIN_NCHANNELS=32 NCCL_MAX_NCHANNELS=32
run NCCL_PROTO=Simple NCCL_MIN_NCHANNELS=32
run NCCL_PROTO=LL128
