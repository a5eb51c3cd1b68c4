#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_pose.py tests/test_gpu_flow.py tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/dev/layer_times.py both split > gpurun_out/layers_sat.txt 2>&1; grep -E "^==|total conv" gpurun_out/layers_sat.txt
timeout 300 python bench.py --no-cpu-baseline 2>/dev/null > gpurun_out/bench_sat.json; python -c "import json; d=json.load(open('gpurun_out/bench_sat.json')); print('bench', d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['achieved'])"
