#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_flow.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/dev/corr_bench.py > gpurun_out/corr_bench.txt 2>&1; cat gpurun_out/corr_bench.txt
timeout 300 python tools/dev/config5_bench.py 1 8 2>&1 | tail -2
