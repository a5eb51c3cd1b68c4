#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -X faulthandler tools/dev/corr_bench.py > gpurun_out/corr_bench.txt 2>&1; cat gpurun_out/corr_bench.txt
