#!/bin/bash
mkdir -p gpurun_out
for m in split fast; do for sk in 0 1 2 3 4 7 8 12 15; do FT_DEV_SKIP=$sk timeout 60 python tools/dev/corr_skip.py $m; done; done > gpurun_out/corr_skip.txt 2>&1
cat gpurun_out/corr_skip.txt
