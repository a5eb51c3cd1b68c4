#!/bin/bash
# re-establish the round baseline on a fresh box: gpu tests, bench, per-layer tables
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; tail -3 gpurun_out/pytest_gpu.txt
timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 300 python tools/dev/layer_times.py both split > gpurun_out/layers_split.txt 2>&1
tail -2 gpurun_out/layers_split.txt
