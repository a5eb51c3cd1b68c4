#!/bin/bash
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
python -c "import json; d=json.load(open('gpurun_out/bench.json')); print('bench', d['value'], d['e2e']['value'], d['ms_per_step'], d['clocks'])"
