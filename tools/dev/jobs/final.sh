#!/bin/bash
# round-end run on one GPU: full GPU suite, smoke, the default bench line (all legs), the reference arm, config 5, the 2-pair shard
T=${1:-r02y}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/${T}_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.txt 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/${T}_smoke.txt
python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference > gpurun_out/${T}_bench_n1_reference_arm.json 2> gpurun_out/${T}_ref.err; echo "ref rc=$?"
python bench.py --config 5 > gpurun_out/${T}_bench_cfg5_n1.json 2> gpurun_out/${T}_cfg5.err; echo "cfg5 rc=$?"
python bench.py --pairs 2 --no-cpu-baseline --no-reference-gpu --no-kernels > gpurun_out/${T}_bench_p2_n1.json 2> gpurun_out/${T}_p2.err; echo "p2 rc=$?"
python - <<PY
import json
for f in ("bench_n1","bench_n1_reference_arm","bench_cfg5_n1","bench_p2_n1"):
    d=json.loads(open("gpurun_out/${T}_%s.json"%f).read().strip().splitlines()[-1])
    c=((d.get("kernels") or {}).get("correlation") or {})
    print(f, round(d["value"],1), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), (d.get("roofline") or {}).get("traffic"), (d.get("parity") or {}).get("ok"), (d.get("cpu_baseline") or {}).get("value"), (d.get("clocks") or {}).get("sm_mhz"), c.get("ms"), c.get("frac"))
PY
