#!/bin/bash
# 8-GPU validation under the driver's launcher: the sharded YouTube-VOS-shaped batch (bench.py --gpus 8) and the DDP training step (config 5).  Output -> gpurun_out/
cd "$(dirname "$0")/.."
N=${N:-8}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench_${N}gpu.log 2>&1
tail -c 1800 gpurun_out/r2_bench_${N}gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/train_step_bench.py --steps 10 --warmup 3 > gpurun_out/r2_train_${N}gpu.log 2>&1
tail -c 1200 gpurun_out/r2_train_${N}gpu.log
