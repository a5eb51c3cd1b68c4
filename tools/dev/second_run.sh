#!/bin/bash
mkdir -p gpurun_out
for t in ops conv pose flow; do
  timeout 1200 python -m pytest tests/test_gpu_$t.py -q -m gpu -s -p no:cacheprovider > gpurun_out/test_$t.log 2>&1; echo "rc=$?" >> gpurun_out/test_$t.log
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/pose_r50_b64_split_launches.csv python tools/dev/pose_once.py 50 64 split 2 > gpurun_out/pose_once_split.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/pose_r50_b64_fast_launches.csv python tools/dev/pose_once.py 50 64 fast 2 > gpurun_out/pose_once_fast.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 57 -c 57 -o gpurun_out/pose_r50_b64_split_full python tools/dev/pose_once.py 50 64 split 2 > gpurun_out/ncu_full.log 2>&1
for t in ops conv pose flow; do echo "== $t"; grep -E "passed|failed|rc=|max-abs" gpurun_out/test_$t.log | tail -12; done
tail -5 gpurun_out/ncu_full.log; ls -la gpurun_out
