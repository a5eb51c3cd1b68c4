#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -q -m gpu -x -p no:cacheprovider > gpurun_out/test_conv.log 2>&1; echo "rc=$?" >> gpurun_out/test_conv.log
tail -30 gpurun_out/test_conv.log
if grep -q "rc=0" gpurun_out/test_conv.log; then
  timeout 900 python -m pytest tests/test_gpu_pose.py tests/test_gpu_flow.py -q -m gpu -x -p no:cacheprovider -s > gpurun_out/test_nets.log 2>&1; echo "rc=$?" >> gpurun_out/test_nets.log
  tail -15 gpurun_out/test_nets.log
  timeout 600 python tools/dev/layer_times.py both split > gpurun_out/layers_split.txt 2>&1
  timeout 600 python tools/dev/layer_times.py both fast > gpurun_out/layers_fast.txt 2>&1
  grep -E "total|==" gpurun_out/layers_split.txt gpurun_out/layers_fast.txt
fi
