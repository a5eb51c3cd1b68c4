#!/bin/bash
# ncu evidence for the current code: launch list of the bench command + full-set capture of every conv launch of one pose forward
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01d_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r01d_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 57 -c 57 -o gpurun_out/r01d_pose_full python tools/dev/pose_once.py 50 64 split 2 > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/r01d_pose_full.ncu-rep --page raw --csv --metrics $M > gpurun_out/r01d_pose_full_raw.csv 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 23 -c 23 -o gpurun_out/r01d_flow_full python tools/dev/flow_once.py 16 split 2 > gpurun_out/ncu_full_flow.log 2>&1
ncu -i gpurun_out/r01d_flow_full.ncu-rep --page raw --csv --metrics $M > gpurun_out/r01d_flow_full_raw.csv 2>&1
tail -3 gpurun_out/ncu_full.log gpurun_out/ncu_full_flow.log; ls -la gpurun_out; head -5 gpurun_out/r01d_pose_full_raw.csv
