#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 python tools/profile_phases.py > gpurun_out/${TAG:-r2}_phases.log 2>&1
grep -A8 "readout_fused CTA0" gpurun_out/${TAG:-r2}_phases.log | head -${LINES_SHOWN:-24}; grep "per call" gpurun_out/${TAG:-r2}_phases.log | head -4
