#!/bin/bash
mkdir -p gpurun_out
for sk in 0 8 7 15; do echo "== skip=$sk"; FT_DEV_SKIP=$sk CB_ONLY="5:128:1,8:128:1,0:128:1,6:128:1" timeout 300 python tools/dev/conv_bench.py split; done > gpurun_out/conv_skip_store.txt 2>&1
cat gpurun_out/conv_skip_store.txt
