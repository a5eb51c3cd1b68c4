#!/bin/bash
# First GPU contact: probe -> tests (no -x, bounded) -> quick bench.  Output lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python tools/dev/tc_probe.py > gpurun_out/probe.log 2>&1; echo "probe rc=$?" >> gpurun_out/probe.log
timeout 900 python -m pytest tests/test_gpu_ops.py -q -m gpu -p no:cacheprovider > gpurun_out/test_ops.log 2>&1; echo "rc=$?" >> gpurun_out/test_ops.log
timeout 900 python -m pytest tests/test_gpu_conv.py -q -m gpu -p no:cacheprovider > gpurun_out/test_conv.log 2>&1; echo "rc=$?" >> gpurun_out/test_conv.log
timeout 900 python -m pytest tests/test_gpu_pose.py -q -m gpu -s -p no:cacheprovider > gpurun_out/test_pose.log 2>&1; echo "rc=$?" >> gpurun_out/test_pose.log
timeout 600 python tools/dev/quick_bench.py > gpurun_out/quick_bench.log 2>&1; echo "rc=$?" >> gpurun_out/quick_bench.log
tail -n 25 gpurun_out/probe.log gpurun_out/test_ops.log gpurun_out/test_conv.log gpurun_out/test_pose.log gpurun_out/quick_bench.log
