#!/bin/bash
# compute-sanitizer over one memorize x4 + readout x3 at the DAVIS shape (N = 2 objects to keep the run short).  Output -> gpurun_out/
cd "$(dirname "$0")/.."
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/run_once.py 2 > gpurun_out/r2_san_$tool.log 2>&1
  tail -4 gpurun_out/r2_san_$tool.log
done
