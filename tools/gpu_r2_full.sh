#!/bin/bash
# full GPU test-suite (no -x) + EM phases.  Output -> gpurun_out/
cd "$(dirname "$0")/.."
TAG=${TAG:-r2full}
rm -f gpurun_out/parity_report.txt
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/${TAG}_tests.log
cp gpurun_out/parity_report.txt gpurun_out/${TAG}_parity_report.txt 2>/dev/null
tail -15 gpurun_out/${TAG}_tests.log
timeout 300 python tools/profile_phases.py > gpurun_out/${TAG}_phases.log 2>&1
grep -v "^  " gpurun_out/${TAG}_phases.log | head -30
sed -n 1,${LINES_SHOWN:-30}p gpurun_out/${TAG}_phases.log
