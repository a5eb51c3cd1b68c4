#!/bin/bash
# split-K reduction rewrite: correctness (conv tests) + autotune logs / layer tables at the bench shapes and at a strong-scaling shard
mkdir -p gpurun_out
python -m pytest tests/test_gpu_conv.py -m gpu -q -x > gpurun_out/r02s_pytest_conv.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02s_pytest_conv.txt
FT_AUTOTUNE_LOG=1 python tools/dev/layer_times.py both split 64 16 > gpurun_out/r02s_layers_b64.txt 2>&1; echo "b64 rc=$?"
FT_AUTOTUNE_LOG=1 python tools/dev/layer_times.py both split 8 2 > gpurun_out/r02s_layers_b8.txt 2>&1; echo "b8 rc=$?"
grep "summed" gpurun_out/r02s_layers_b64.txt gpurun_out/r02s_layers_b8.txt
