#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_ops.py -q -m gpu -x -p no:cacheprovider > gpurun_out/test_conv.log 2>&1; echo "rc=$?" >> gpurun_out/test_conv.log
tail -5 gpurun_out/test_conv.log
if grep -q "rc=0" gpurun_out/test_conv.log; then
  timeout 900 python -m pytest tests/test_gpu_pose.py tests/test_gpu_flow.py -q -m gpu -x -p no:cacheprovider -s > gpurun_out/test_nets.log 2>&1; echo "rc=$?" >> gpurun_out/test_nets.log
  tail -4 gpurun_out/test_nets.log
  timeout 600 python tools/dev/layer_times.py both split > gpurun_out/layers_split.txt 2>&1
  grep -E "total|==" gpurun_out/layers_split.txt
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
  python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['kernels']['other_kernels_ms_per_step'])"
fi
