#!/bin/bash
# Round-2 final evidence: ncu launch list of the bench command (cuDNN autotuning off: its trial launches would flood the list), one
# --set full capture per hot kernel, hot-path launch list, phase stamps, micro-benchmark sweep with the eager-CUDA reference columns,
# fusion-layer bench, compute-sanitizer (N = 2).  Output -> gpurun_out/r2z_*
cd "$(dirname "$0")/.."
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2z_launches.csv \
    env SWEM_CUDNN_BENCHMARK=0 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-batch > gpurun_out/ncu_bench.log 2>&1
for k in em_res_kernel readout_topl_kernel fusion_conv_glu_kernel fusion_act_images_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/r2z_$k \
      python tools/run_once.py > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log
done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2z_hot_launches.csv python tools/run_once.py > /dev/null 2>&1
timeout 300 python tools/profile_phases.py > gpurun_out/r2z_phases.log 2>&1
timeout 600 python tools/microbench.py --reps 40 --eager > gpurun_out/r2z_microbench.txt 2>&1
tail -3 gpurun_out/r2z_microbench.txt
timeout 200 python tools/fusion_bench.py > gpurun_out/r2z_fusion_bench.txt 2>&1
cat gpurun_out/r2z_fusion_bench.txt
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/run_once.py 2 > gpurun_out/r2z_san_$tool.log 2>&1
  tail -3 gpurun_out/r2z_san_$tool.log
done
ls -la gpurun_out | grep r2z_ | tail -20
