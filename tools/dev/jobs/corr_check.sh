python -m pytest tests/test_gpu_ops.py tests/test_gpu_flow.py tests/test_gpu_baseline_sizes.py -m gpu -q -x 2>&1 | tail -2
python tools/dev/corr_bench.py 2>&1 | grep "\"ms\|ms_\|frac"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__m_xbar2l1tex_read_bytes.sum,l1tex__m_l1tex2xbar_write_bytes.sum,sm__cycles_active.avg,launch__registers_per_thread,launch__grid_size,smsp__inst_executed.sum"
timeout 600 ncu --set full --clock-control none -k regex:corr_tc_kernel -c 4 -o gpurun_out/corr_full python tools/dev/corr_bench.py > gpurun_out/r02_ncu_corr.log 2>&1
ncu -i gpurun_out/corr_full.ncu-rep --page raw --csv --metrics $M > gpurun_out/r02_corr_raw.csv 2>&1
rm -f gpurun_out/corr_full.ncu-rep; wc -l gpurun_out/r02_corr_raw.csv
