#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/test_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/test_gpu.log
tail -8 gpurun_out/test_gpu.log
timeout 600 python tools/dev/layer_times.py both split > gpurun_out/layers_split.txt 2>&1
grep -E "total|==|^ +[0-2] " gpurun_out/layers_split.txt | cut -c1-180
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['kernels']['other_kernels_ms_per_step'])"
