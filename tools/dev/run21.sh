#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; tail -3 gpurun_out/pytest_gpu.txt
timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print(d['value'], d['e2e']['value'], d['roofline']['achieved'], json.dumps(d['kernels']['correlation'], indent=1))
PY
