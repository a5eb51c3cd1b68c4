#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py -q -m gpu -x -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/test_pipe.log; cat gpurun_out/test_pipe.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['ms_per_step'], d['e2e'])"
for bn in 64 256; do
FT_DEV_BLOCK_N=$bn timeout 600 python tools/dev/layer_times.py both split > gpurun_out/layers_split_bn$bn.txt 2>&1
grep -E "total|==" gpurun_out/layers_split_bn$bn.txt | cut -c1-180
done
timeout 600 python tools/dev/layer_times.py both split > gpurun_out/layers_split.txt 2>&1
