#!/bin/bash
# full GPU suite + smoke + a short bench line (no CPU / reference legs); logs under gpurun_out/ with the tag given as $1
T=${1:-r02x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${T}_smoke.txt
python bench.py --no-cpu-baseline --no-reference-gpu --no-kernels > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"]["sm_mhz"], d["parity"]["ok"] if d.get("parity") else None, d["roofline"]["kernel_ms_per_step"], d["kernels"]["other_kernels_ms_per_step"])
PY
