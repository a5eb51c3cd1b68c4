#!/bin/bash
# round-1 GPU pass: parity tests, smoke, bench, per-layer times, ncu launch list + full capture of a few conv launches
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 1500 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/test_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/test_gpu.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 python tools/dev/layer_times.py both split > gpurun_out/layers_split.txt 2>&1
timeout 600 python tools/dev/layer_times.py both fast > gpurun_out/layers_fast.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 62 -c 6 -o gpurun_out/pose_split_l1 python tools/dev/pose_once.py 50 64 split 2 > gpurun_out/ncu_full1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 108 -c 6 -o gpurun_out/pose_split_tail python tools/dev/pose_once.py 50 64 split 2 > gpurun_out/ncu_full2.log 2>&1
for f in gpurun_out/*.ncu-rep; do ncu -i $f --page raw --csv > ${f%.ncu-rep}_raw.csv 2>/dev/null; done
tail -3 gpurun_out/test_gpu.log; tail -3 gpurun_out/smoke.log; cat gpurun_out/bench.json | head -c 3000; tail -2 gpurun_out/bench.err
du -sh gpurun_out
