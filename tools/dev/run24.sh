#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/dev/config5_bench.py 1 8 > gpurun_out/config5.txt 2>&1; tail -4 gpurun_out/config5.txt
