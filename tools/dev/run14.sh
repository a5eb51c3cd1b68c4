#!/bin/bash
mkdir -p gpurun_out
export CB_ONLY="0:128:0,0:128:1,0:256:1,3:128:0,3:256:1"
for sk in 0 1 2 3 4 5 6 7; do echo "== skip=$sk"; FT_DEV_SKIP=$sk FT_DEV_CHUNK=1000 timeout 120 python tools/dev/conv_bench.py split; done > gpurun_out/skip_split.txt 2>&1
for sk in 0 1 2 4 7; do echo "== skip=$sk"; FT_DEV_SKIP=$sk timeout 120 python tools/dev/conv_bench.py fast; done > gpurun_out/skip_fast.txt 2>&1
cat gpurun_out/skip_split.txt gpurun_out/skip_fast.txt
