#!/bin/bash
mkdir -p gpurun_out
export CB_ONLY="0:128:0,0:128:1,0:256:1" CB_REPS=1
timeout 600 ncu --set full --clock-control none -k regex:conv_tc -s 9 -c 3 -o gpurun_out/pair_split python tools/dev/conv_bench.py split > gpurun_out/ncu1.log 2>&1
export CB_ONLY="0:128:0,0:256:0"
timeout 600 ncu --set full --clock-control none -k regex:conv_tc -s 6 -c 2 -o gpurun_out/single_fast python tools/dev/conv_bench.py fast > gpurun_out/ncu2.log 2>&1
for f in gpurun_out/*.ncu-rep; do ncu -i $f --page raw --csv > ${f%.ncu-rep}_raw.csv 2>/dev/null; done
tail -3 gpurun_out/ncu1.log; ls -la gpurun_out
