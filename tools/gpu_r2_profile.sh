#!/bin/bash
# Round-2 ncu evidence: launch list of the bench command (cuDNN autotuning off: its trial launches would flood the list), one --set full
# capture per hot kernel, the hot-path launch list, phase stamps.  Output -> gpurun_out/
cd "$(dirname "$0")/.."
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches.csv \
    env SWEM_CUDNN_BENCHMARK=0 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-batch > gpurun_out/ncu_bench.log 2>&1
for k in em_res_kernel readout_topl_kernel readout_prep_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/r2_$k \
      python tools/run_once.py > gpurun_out/ncu_$k.log 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_hot_launches.csv python tools/run_once.py > /dev/null 2>&1
timeout 300 python tools/profile_phases.py > gpurun_out/r2_phases.log 2>&1
ls -la gpurun_out | grep r2_ | tail -12
