#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; tail -3 gpurun_out/pytest_gpu.txt
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
python -c "import json; d=json.load(open('gpurun_out/bench.json')); print('bench', d['value'], d['e2e'], d['ms_per_step'], d['roofline']['frac'], d['cpu_baseline']['value'], d['kernels']['correlation']['ms'], d['gpu_launches_per_step'])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null; cut -c1-300 gpurun_out/bench_reference.json
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
