#!/bin/bash
# A/B of library builds on ONE box, interleaved (box-to-box variance is several per cent): build/ab/lib_<name>.so via FT_LIB_PATH
# usage: ab.sh <pairs> <name> <name> ...
mkdir -p gpurun_out
P=$1; shift
for round in 1 2; do
  for name in "$@"; do
    FT_LIB_PATH=$PWD/build/ab/lib_$name.so python bench.py --pairs $P --steps 40 --warmup 3 --no-kernels --no-reference-gpu --no-cpu-baseline --no-parity > gpurun_out/ab_${name}_p${P}_$round.json 2>/dev/null
    python -c "
import json; d=json.load(open('gpurun_out/ab_${name}_p${P}_$round.json')); print('$name pairs=$P round $round', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), round(d['roofline']['kernel_ms_per_step'],3), d['clocks']['sm_mhz'])"
  done
done
