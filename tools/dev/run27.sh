#!/bin/bash
mkdir -p gpurun_out
for sk in 0 1 2 4 8 7 15; do echo "== skip=$sk"; FT_DEV_SKIP=$sk CB_ONLY="11:64:1,11:64:0,12:64:1,10:64:1" timeout 300 python tools/dev/conv_bench.py split; done > gpurun_out/conv_skip_stem.txt 2>&1
cat gpurun_out/conv_skip_stem.txt
