#!/bin/bash
# Round-2 GPU round trip A: full GPU test-suite + the new bench line.  Output -> gpurun_out/
cd "$(dirname "$0")/.."
rm -f gpurun_out/parity_report.txt
python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r2a_tests.log
cp gpurun_out/parity_report.txt gpurun_out/r2a_parity_report.txt 2>/dev/null
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench.log 2>&1
tail -c 1200 gpurun_out/r2a_tests.log
tail -c 3000 gpurun_out/r2a_bench.log
