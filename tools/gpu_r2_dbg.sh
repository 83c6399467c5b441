#!/bin/bash
# GPU tests (no -x) + phase stamps of both kernels, then the readout's measurement switches (SWEM_RO_DBG)
cd "$(dirname "$0")/.."
TAG=${TAG:-r2dbg}
rm -f gpurun_out/parity_report.txt
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/${TAG}_tests.log
tail -6 gpurun_out/${TAG}_tests.log
timeout 300 python tools/profile_phases.py > gpurun_out/${TAG}_phases.log 2>&1
sed -n 1,75p gpurun_out/${TAG}_phases.log
grep "per call" gpurun_out/${TAG}_phases.log
for d in 1 2 3; do
  echo "--- SWEM_RO_DBG=$d"
  SWEM_RO_DBG=$d timeout 300 python tools/profile_phases.py > gpurun_out/${TAG}_phases_dbg$d.log 2>&1
  grep -A8 "pixel-major output" gpurun_out/${TAG}_phases_dbg$d.log | head -9; grep "per call" gpurun_out/${TAG}_phases_dbg$d.log | head -3
done
