#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "correlation" 2>&1 | tail -3
( for pt in 1 3 5; do FT_DEV_CORR_PARTS=$pt timeout 60 python tools/dev/corr_skip.py split; done
for sk in 3 4 8 15; do FT_DEV_SKIP=$sk timeout 60 python tools/dev/corr_skip.py split; done
timeout 60 python tools/dev/corr_skip.py fast ) > gpurun_out/corr_v2u.txt 2>&1
cat gpurun_out/corr_v2u.txt
