#!/bin/bash
mkdir -p gpurun_out
( for b in 8 16 32; do for pt in 1 2 3 4 5 6; do CORR_B=$b FT_DEV_CORR_PARTS=$pt timeout 60 python tools/dev/corr_skip.py split; done; done ) > gpurun_out/corr_parts.txt 2>&1
cat gpurun_out/corr_parts.txt
