#!/bin/bash
# Round deliverables in one GPU call: full GPU test-suite, smoke, bench (both arms), ncu launch list + full captures,
# phase stamps, stage-kernel and EM/readout micro-benchmarks.  Output -> gpurun_out/ (copied into profiles/ by
# tools/summarize_profiles.py + by hand).
cd "$(dirname "$0")/.."
rm -f gpurun_out/parity_report.txt
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r1_tests.log; cat gpurun_out/r1_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r1_bench_ref.log 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/r1_bench.log 2>&1
python tools/profile_phases.py > gpurun_out/r1_phases_v2.log 2>&1
python tools/stage_bench.py > gpurun_out/r1_stage_bench.txt 2>&1
bash tools/gpu_profile.sh > /dev/null 2>&1
python - <<'PY'
import json
for f in ('gpurun_out/r1_bench.log', 'gpurun_out/r1_bench_ref.log'):
    try:
        l=[x for x in open(f) if x.startswith('{')][-1]; d=json.loads(l)
        print(f, 'value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'launches', d.get('gpu_launches'), 'roof', d.get('roofline',{}).get('frac'), 'cpu', d.get('cpu_baseline',{}).get('value'), 'fp32_convs', d.get('fp32_convs',{}).get('value'))
    except Exception as e:
        print(f, 'ERR', e); print(open(f).read()[-1500:])
PY
