#!/bin/bash
# communication probe + EM phase profile with sub-stamps
cd "$(dirname "$0")/.."
timeout 120 tools/comm_probe > gpurun_out/r2_comm_probe.txt 2>&1
cat gpurun_out/r2_comm_probe.txt
timeout 300 python tools/profile_phases.py > gpurun_out/r2em_phases.log 2>&1
sed -n 1,60p gpurun_out/r2em_phases.log
