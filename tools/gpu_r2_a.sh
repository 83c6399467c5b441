#!/bin/bash
# Round-2 GPU round trip A: full GPU test-suite + the new bench line.  Output -> gpurun_out/
cd "$(dirname "$0")/.."
rm -f gpurun_out/parity_report.txt
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2a_tests.log
cp gpurun_out/parity_report.txt gpurun_out/r2a_parity_report.txt 2>/dev/null
timeout 900 python bench.py --steps 20 --warmup 3 ${BENCH_ARGS} > gpurun_out/r2a_bench.log 2>&1
tail -c 1500 gpurun_out/r2a_tests.log
tail -c 3000 gpurun_out/r2a_bench.log
