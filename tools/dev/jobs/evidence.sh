#!/bin/bash
# Round-end ncu evidence for the code as built (run on the GPU box through gpurun):
#   1. launch list (gpu__time_duration) of the bench.py command            -> gpurun_out/r02_bench_launches.csv
#   2. ncu --set full over every conv launch of one pose / flow forward     -> gpurun_out/r02_{pose,flow}_full_raw.csv
#   3. ncu --set full of corr_pair_kernel at BASELINE config 3              -> gpurun_out/r02_corr_raw.csv
# tools/dev/ncu_summary.py (run here, no GPU needed) turns 2-3 into profiles/r02_conv_traffic.json / r02_corr_ncu.json.
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__m_xbar2l1tex_read_bytes.sum,l1tex__m_l1tex2xbar_write_bytes.sum,sm__cycles_active.avg,launch__registers_per_thread,launch__grid_size,smsp__inst_executed.sum"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-gpu --no-kernels --no-parity > gpurun_out/r02_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:conv_tc -o gpurun_out/pose_full python tools/dev/pose_once.py 50 64 split 2 > gpurun_out/r02_ncu_pose.log 2>&1
ncu -i gpurun_out/pose_full.ncu-rep --page raw --csv --metrics $M > gpurun_out/r02_pose_full_raw.csv 2>&1
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:conv_tc -o gpurun_out/flow_full python tools/dev/flow_once.py 16 split 2 > gpurun_out/r02_ncu_flow.log 2>&1
ncu -i gpurun_out/flow_full.ncu-rep --page raw --csv --metrics $M > gpurun_out/r02_flow_full_raw.csv 2>&1
timeout 600 ncu --set full --clock-control none -k regex:corr_pair -c 4 -o gpurun_out/corr_full python tools/dev/corr_bench.py > gpurun_out/r02_ncu_corr.log 2>&1
ncu -i gpurun_out/corr_full.ncu-rep --page raw --csv --metrics $M > gpurun_out/r02_corr_raw.csv 2>&1
rm -f gpurun_out/pose_full.ncu-rep gpurun_out/flow_full.ncu-rep gpurun_out/corr_full.ncu-rep
tail -2 gpurun_out/r02_ncu_pose.log gpurun_out/r02_ncu_flow.log gpurun_out/r02_ncu_corr.log; wc -l gpurun_out/r02_*raw.csv gpurun_out/r02_bench_launches.csv
