#!/bin/bash
mkdir -p gpurun_out
for ch in 4 1000; do FT_DEV_CHUNK=$ch timeout 300 python tools/dev/conv_bench.py split; done > gpurun_out/conv_bench.txt 2>&1
timeout 300 python tools/dev/conv_bench.py fast >> gpurun_out/conv_bench.txt 2>&1
grep -v "BN=64" gpurun_out/conv_bench.txt
