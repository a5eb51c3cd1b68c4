#!/bin/bash
mkdir -p gpurun_out
CB_ONLY="5:128:0,5:128:1,5:256:0,5:256:1,6:128:0,6:128:1,6:256:1,7:128:1,7:256:1,7:256:0,8:128:1,8:256:1,8:256:0,9:128:1,9:256:1,10:64:0,10:64:1" timeout 300 python tools/dev/conv_bench.py split > gpurun_out/conv_tilings_1x1.txt 2>&1
cat gpurun_out/conv_tilings_1x1.txt
