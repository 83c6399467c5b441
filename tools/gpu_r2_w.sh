#!/bin/bash
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -q -x -k "engine or bench_configuration or free_running or pipelined or graph" 2>&1 | tail -3
python bench.py --no-cpu-baseline --no-batch > gpurun_out/r2w_bench.log 2> gpurun_out/r2w_bench.err
python - <<PY
import json
l=[x for x in open("gpurun_out/r2w_bench.log") if x.startswith("{")][-1]; d=json.loads(l); r=d["roofline"]
print("fps", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "em_us", round(r["em_us"],1), "read_us", round(r["readout_us"],1), "launches", d["gpu_launches"])
PY
