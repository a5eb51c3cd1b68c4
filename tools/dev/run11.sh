#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv.py -q -m gpu -x -p no:cacheprovider -k "cta_pair" > gpurun_out/test_pair.log 2>&1; echo "rc=$?" >> gpurun_out/test_pair.log
tail -25 gpurun_out/test_pair.log | cut -c1-300
if grep -q "rc=0" gpurun_out/test_pair.log; then
  timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -5 > gpurun_out/test_gpu.log; cat gpurun_out/test_gpu.log
  timeout 600 python tools/dev/layer_times.py both split > gpurun_out/layers_split.txt 2>&1
  grep -E "total|==" gpurun_out/layers_split.txt | cut -c1-180
  FT_DEV_PAIR=1 timeout 600 python tools/dev/layer_times.py both split > gpurun_out/layers_split_pair1.txt 2>&1
  grep -E "total|==" gpurun_out/layers_split_pair1.txt | cut -c1-180
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
  python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'])"
fi
