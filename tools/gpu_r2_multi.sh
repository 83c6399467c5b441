#!/bin/bash
# multi-GPU bench (the driver's launcher) at N GPUs: the sharded YouTube-VOS-shaped batch.  Output -> gpurun_out/
cd "$(dirname "$0")/.."
N=${N:-2}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 ${BENCH_ARGS} > gpurun_out/r2_bench_${N}gpu.log 2>&1
tail -c 3500 gpurun_out/r2_bench_${N}gpu.log
