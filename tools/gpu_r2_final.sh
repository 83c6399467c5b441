#!/bin/bash
# full GPU test-suite + default bench (N = 1).  Output -> gpurun_out/${TAG}_*
cd "$(dirname "$0")/.."
TAG=${TAG:-r2f}
rm -f gpurun_out/parity_report.txt
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/${TAG}_tests.log
cp gpurun_out/parity_report.txt gpurun_out/${TAG}_parity_report.txt 2>/dev/null
tail -6 gpurun_out/${TAG}_tests.log
SECONDS=0; python bench.py > gpurun_out/${TAG}_bench.log 2> gpurun_out/${TAG}_bench.err
echo "bench wall ${SECONDS}s"
python - <<PY
import json
try:
    l=[x for x in open('gpurun_out/${TAG}_bench.log') if x.startswith('{')][-1]; d=json.loads(l)
    print('fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'em_us', round(d['roofline'].get('em_us',0),1), 'read_us', round(d['roofline'].get('readout_us',0),1),
          'frac', round(d['roofline']['frac'],4), 'parity', d.get('parity',{}).get('min_frame_agreement'), 'launches', d.get('gpu_launches'))
except Exception as e:
    print('ERR', e); print(open('gpurun_out/${TAG}_bench.err').read()[-3000:])
PY
