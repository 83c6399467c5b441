#!/bin/bash
# EM pair kernel bring-up: fused-family parity tests, phase profile v2 vs v1, bench
cd "$(dirname "$0")/.."
rm -f gpurun_out/parity_report.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused or reproducible or free_running or engine or ytvos" 2>&1 | tail -15 > gpurun_out/r1_tests_v2.log
tail -5 gpurun_out/r1_tests_v2.log
timeout 300 python tools/profile_phases.py > gpurun_out/r1_phases_v2.log 2>&1
grep -v "^  " gpurun_out/r1_phases_v2.log; head -40 gpurun_out/r1_phases_v2.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r1_bench_v2.log 2>&1
python - <<'PY'
import json
try:
    l=[x for x in open('gpurun_out/r1_bench_v2.log') if x.startswith('{')][-1]; d=json.loads(l)
    print('fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'em_us', round(d['roofline']['em_us'],1), 'read_us', round(d['roofline']['readout_us'],1))
except Exception as e:
    print('ERR', e); print(open('gpurun_out/r1_bench_v2.log').read()[-2000:])
PY
