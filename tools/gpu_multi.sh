#!/bin/bash
# multi-GPU validation (N = number of visible GPUs): bench.py, YTVOS batch, DDP training step
cd "$(dirname "$0")/.."
N=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r1_bench_${N}gpu.log 2>&1; tail -c 700 gpurun_out/r1_bench_${N}gpu.log; echo
timeout 600 $TR tools/ytvos_bench.py --sequences 8 --max-frames 12 > gpurun_out/r1_ytvos_${N}gpu.log 2>&1; tail -c 500 gpurun_out/r1_ytvos_${N}gpu.log; echo
timeout 300 python tools/ytvos_bench.py --sequences 8 --max-frames 12 > gpurun_out/r1_ytvos_1of${N}gpu.log 2>&1; tail -c 500 gpurun_out/r1_ytvos_1of${N}gpu.log; echo
timeout 600 $TR tools/train_step_bench.py --steps 5 --warmup 2 > gpurun_out/r1_train_${N}gpu.log 2>&1; tail -c 500 gpurun_out/r1_train_${N}gpu.log; echo
