#!/bin/bash
# One GPU round trip: sanitizer on the flaky test, full GPU test-suite, bench variants.  Output -> gpurun_out/
cd "$(dirname "$0")/.."
rm -f gpurun_out/parity_report.txt
T='tests/test_gpu_parity.py::test_golden_sequences_teacher_forced'
for tool in initcheck racecheck memcheck; do
  timeout 400 compute-sanitizer --tool $tool --kernel-regex kns=swem python -m pytest "$T" -x -q -k "core_small and generic" > gpurun_out/san_$tool.log 2>&1
done
rm -f gpurun_out/parity_report.txt
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r1_tests.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r1_bench.log 2>&1
SWEM_FUSED_CONV=0 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r1_bench_nofusedconv.log 2>&1
SWEM_ENGINE=0 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r1_bench_noengine.log 2>&1
tail -c 1500 gpurun_out/r1_tests.log
for f in gpurun_out/san_*.log; do echo "== $f"; grep -c "=========" $f; tail -4 $f; done
for f in gpurun_out/r1_bench.log gpurun_out/r1_bench_nofusedconv.log gpurun_out/r1_bench_noengine.log; do python - "$f" <<'PY'
import json,sys
try:
    l=[x for x in open(sys.argv[1]) if x.startswith('{')][-1]; d=json.loads(l)
    print(sys.argv[1], 'fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'em_us', round(d['roofline']['em_us'],1), 'read_us', round(d['roofline']['readout_us'],1), 'eager_ms', round(d['config']['eager_ms_per_step'],2))
except Exception as e:
    print(sys.argv[1], 'ERR', e); print(open(sys.argv[1]).read()[-1500:])
PY
done
