#!/bin/bash
mkdir -p gpurun_out
python tools/dev/tma_probe.py > gpurun_out/tma_probe.log 2>&1; cat gpurun_out/tma_probe.log
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_flow.py -q -m gpu -x -p no:cacheprovider > gpurun_out/test_conv.log 2>&1; echo "rc=$?" >> gpurun_out/test_conv.log
tail -5 gpurun_out/test_conv.log
timeout 600 python tools/dev/layer_times.py flow split > gpurun_out/layers_split.txt 2>&1
cat gpurun_out/layers_split.txt | cut -c1-180
