#!/bin/bash
# Round-end ncu evidence for the code as built (run on the GPU box through gpurun):
#   1. launch list (gpu__time_duration + DRAM bytes) of exactly the timed device steps of bench.py (--profile-range):
#      at the bench shape (16 pairs + 64 crops)                              -> gpurun_out/r02_bench_launches.csv
#      and at a strong-scaling shard (2 pairs + 8 crops = 1/8 of the batch)   -> gpurun_out/r02_bench_launches_p2.csv
#   2. ncu --set full over every conv launch of one pose / flow forward     -> gpurun_out/r02_{pose,flow}_full_raw.csv
#   3. ncu --set full of corr_tc_kernel (the default correlation kernel) at BASELINE config 3              -> gpurun_out/r02_corr_raw.csv
#   4. compute-sanitizer memcheck over the split-K conv tests and the elementwise / tracking kernels -> gpurun_out/r02_sanitizer.txt
# tools/dev/ncu_summary.py (run here, no GPU needed) turns 2-3 into profiles/r02_conv_traffic.json / r02_corr_ncu.json.
# usage: evidence.sh [all|lists]   (lists = only step 1)
MODE=${1:-all}
mkdir -p gpurun_out
# tile choices are measured ONCE outside the profiler (ncu perturbs the autotuner's timings) and loaded by every run below
export FT_AUTOTUNE_CACHE=$PWD/gpurun_out/r02_tuned.json
rm -f $FT_AUTOTUNE_CACHE
python bench.py --no-cpu-baseline --no-reference-gpu --no-kernels --no-parity --steps 2 --warmup 3 > /dev/null 2>&1
python bench.py --no-cpu-baseline --no-reference-gpu --no-kernels --no-parity --steps 2 --warmup 3 --pairs 2 > /dev/null 2>&1
[ "$MODE" = all ] && { python tools/dev/pose_once.py 50 64 split 1 > /dev/null 2>&1; python tools/dev/flow_once.py 16 split 1 > /dev/null 2>&1; }
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__m_xbar2l1tex_read_bytes.sum,l1tex__m_l1tex2xbar_write_bytes.sum,sm__cycles_active.avg,launch__registers_per_thread,launch__grid_size,smsp__inst_executed.sum"
B="--no-cpu-baseline --no-reference-gpu --no-kernels --no-parity --profile-range --steps 2 --warmup 3"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py $B > gpurun_out/r02_bench_under_ncu.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_bench_launches_p2.csv python bench.py $B --pairs 2 > gpurun_out/r02_bench_under_ncu_p2.log 2>&1
wc -l gpurun_out/r02_bench_launches*.csv
[ "$MODE" = all ] || exit 0
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:conv_tc -o gpurun_out/pose_full python tools/dev/pose_once.py 50 64 split 2 > gpurun_out/r02_ncu_pose.log 2>&1
ncu -i gpurun_out/pose_full.ncu-rep --page raw --csv --metrics $M > gpurun_out/r02_pose_full_raw.csv 2>&1
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:conv_tc -o gpurun_out/flow_full python tools/dev/flow_once.py 16 split 2 > gpurun_out/r02_ncu_flow.log 2>&1
ncu -i gpurun_out/flow_full.ncu-rep --page raw --csv --metrics $M > gpurun_out/r02_flow_full_raw.csv 2>&1
timeout 600 ncu --set full --clock-control none -k regex:corr_tc_kernel -c 4 -o gpurun_out/corr_full python tools/dev/corr_bench.py > gpurun_out/r02_ncu_corr.log 2>&1
ncu -i gpurun_out/corr_full.ncu-rep --page raw --csv --metrics $M > gpurun_out/r02_corr_raw.csv 2>&1
rm -f gpurun_out/pose_full.ncu-rep gpurun_out/flow_full.ncu-rep gpurun_out/corr_full.ncu-rep
FT_AUTOTUNE=0 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_conv.py tests/test_gpu_track.py tests/test_gpu_ops.py -m gpu -q -x -k "splitk or pack_pair or maxpool or upsample or flow_head or crop or nms or decode or boxprop" > gpurun_out/r02_sanitizer.txt 2>&1; echo "sanitizer rc=$?"
tail -n 4 gpurun_out/r02_sanitizer.txt
for f in gpurun_out/r02_ncu_pose.log gpurun_out/r02_ncu_flow.log gpurun_out/r02_ncu_corr.log; do tail -n 2 $f; done; wc -l gpurun_out/r02_*raw.csv gpurun_out/r02_bench_launches*.csv
