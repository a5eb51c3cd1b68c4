#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:corr_tc -s 3 -c 1 -o gpurun_out/corr_v2 python tools/dev/corr_skip.py split > gpurun_out/ncu_corr.log 2>&1
tail -3 gpurun_out/ncu_corr.log
ncu -i gpurun_out/corr_v2.ncu-rep --page details --csv > gpurun_out/corr_v2_details.csv 2>&1
ncu -i gpurun_out/corr_v2.ncu-rep --page raw --csv > gpurun_out/corr_v2_raw.csv 2>&1
ncu -i gpurun_out/corr_v2.ncu-rep --page source --csv > gpurun_out/corr_v2_source.csv 2>&1
ls -la gpurun_out/
