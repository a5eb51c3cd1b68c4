#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_pose.py tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/dev/layer_times.py pose split > gpurun_out/layers_s2d.txt 2>&1; grep -E "^==|total conv" gpurun_out/layers_s2d.txt; sed -n 2,5p gpurun_out/layers_s2d.txt | cut -c1-160
timeout 300 python bench.py --no-cpu-baseline 2>/dev/null > gpurun_out/bench_s2d.json; python -c "import json; d=json.load(open('gpurun_out/bench_s2d.json')); print('bench', d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['achieved'])"
