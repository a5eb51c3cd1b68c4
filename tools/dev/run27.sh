#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python tools/dev/conv_bench.py split; FT_DEV_CHUNK=1000 timeout 300 python tools/dev/conv_bench.py split ) > gpurun_out/conv_tilings_uniform.txt 2>&1
cat gpurun_out/conv_tilings_uniform.txt
