#!/bin/bash
# Round-2 EM bring-up: EM parity tests on the V-resident kernel, phase profile, A/B against em_pair_kernel.  Output -> gpurun_out/
cd "$(dirname "$0")/.."
TAG=${TAG:-r2em}
rm -f gpurun_out/parity_report.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "${TESTS:-fused or golden or iteration or responsibilities or encoder}" 2>&1 | tail -25 > gpurun_out/${TAG}_tests.log
cp gpurun_out/parity_report.txt gpurun_out/${TAG}_parity_report.txt 2>/dev/null
tail -8 gpurun_out/${TAG}_tests.log
timeout 300 python tools/profile_phases.py > gpurun_out/${TAG}_phases.log 2>&1
for d in ${DBGS:-}; do SWEM_EM_DBG=$d timeout 300 python tools/profile_phases.py > gpurun_out/${TAG}_phases_dbg$d.log 2>&1; done
grep -v "^  " gpurun_out/${TAG}_phases.log | head -30
sed -n 1,${LINES_SHOWN:-60}p gpurun_out/${TAG}_phases.log
for d in ${DBGS:-}; do echo "== dbg $d"; grep -E "memorize|total" gpurun_out/${TAG}_phases_dbg$d.log | head -4; sed -n 10,22p gpurun_out/${TAG}_phases_dbg$d.log; done
