#!/bin/bash
set -x
O=gpurun_out/r2i
mkdir -p $O
python -c "import glass_b200.build as b; print(b.build())" > $O/build.log 2>&1
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dp_check.py em_user_shaped > $O/dp_check.json 2> $O/dp_check.err
echo "rc $?" >> $O/dp_check.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 10 > $O/bench_n2_symm.json 2> $O/bench_n2_symm.err
GLASS_B200_DP=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 100 --warmup 10 > $O/bench_n2_nccl.json 2> $O/bench_n2_nccl.err
timeout 600 python bench.py --gpus 1 --steps 100 --warmup 10 --no-cpu-baseline --no-other-configs --no-gpu-eager-baseline > $O/bench_n1.json 2> $O/bench_n1.err
ls -la $O
