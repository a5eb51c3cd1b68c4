#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_pose.py -m gpu -x -q -s -k "384x288" 2>&1 | tail -4
timeout 300 python tools/dev/config5_bench.py 1 8 > gpurun_out/config5.txt 2>&1; tail -2 gpurun_out/config5.txt
