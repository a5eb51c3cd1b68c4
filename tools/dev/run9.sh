#!/bin/bash
mkdir -p gpurun_out
python tools/dev/acc_probe.py 2>&1 | grep "tc" | grep -v positive > gpurun_out/acc_probe.log; 
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -s 2>&1 | grep -E "max-abs|passed|failed|FAILED|Error" > gpurun_out/test_gpu.log; cat gpurun_out/test_gpu.log
timeout 600 python tools/dev/layer_times.py both split > gpurun_out/layers_split.txt 2>&1
grep -E "total|==" gpurun_out/layers_split.txt | cut -c1-180
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['kernels']['other_kernels_ms_per_step'])"
