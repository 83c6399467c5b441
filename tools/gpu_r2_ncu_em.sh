#!/bin/bash
# EM tests + phases + one ncu --set full capture of em_res_kernel (source-level stalls).  Output -> gpurun_out/
cd "$(dirname "$0")/.."
LINES_SHOWN=${LINES_SHOWN:-70} TESTS="${TESTS:-fused or golden or iteration or responsibilities or encoder}" bash tools/gpu_r2_em.sh
ncu --set full --clock-control none --import-source on -k regex:em_res_kernel -s 2 -c 1 -f -o gpurun_out/r2_em_res_kernel \
    python tools/run_once.py > gpurun_out/ncu_em_res_kernel.log 2>&1
tail -3 gpurun_out/ncu_em_res_kernel.log
ls -la gpurun_out/r2_em_res_kernel.ncu-rep
