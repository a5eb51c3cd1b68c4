#!/bin/bash
# ncu evidence for the current code: launch list of the bench command + full-set capture of every conv launch of one pose / flow forward
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__ops_path_tensor_src_fp16_dst_fp32.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum,sm__cycles_active.avg,launch__registers_per_thread,launch__grid_size"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r01j_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r01j_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:conv_tc -s 57 -c 57 -o gpurun_out/pose_full python tools/dev/pose_once.py 50 64 split 2 > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/pose_full.ncu-rep --page raw --csv --metrics $M > gpurun_out/r01j_pose_full_raw.csv 2>&1
timeout 900 ncu --set full --clock-control none -k regex:conv_tc -s 19 -c 19 -o gpurun_out/flow_full python tools/dev/flow_once.py 16 split 2 > gpurun_out/ncu_full_flow.log 2>&1
ncu -i gpurun_out/flow_full.ncu-rep --page raw --csv --metrics $M > gpurun_out/r01j_flow_full_raw.csv 2>&1
ls -la gpurun_out/*.ncu-rep
rm -f gpurun_out/pose_full.ncu-rep gpurun_out/flow_full.ncu-rep
# one capture with source of the largest pose layer (deconv.6: 57th conv launch of a forward = index 56)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 113 -c 1 -o gpurun_out/r01j_deconv6 python tools/dev/pose_once.py 50 64 split 2 > gpurun_out/ncu_deconv6.log 2>&1
ncu -i gpurun_out/r01j_deconv6.ncu-rep --page details --csv > gpurun_out/r01j_deconv6_details.csv 2>&1
ncu -i gpurun_out/r01j_deconv6.ncu-rep --page source --csv > gpurun_out/r01j_deconv6_source.csv 2>&1
tail -2 gpurun_out/ncu_full.log gpurun_out/ncu_full_flow.log gpurun_out/ncu_deconv6.log; du -sh gpurun_out; head -4 gpurun_out/r01j_pose_full_raw.csv | cut -c1-400
