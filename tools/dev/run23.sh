#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -x -q -k "reverse_tile" 2>&1 | tail -3
for o in fixed alternate; do echo "== FT_TILE_ORDER=$o"; FT_TILE_ORDER=$o timeout 300 python tools/dev/layer_times.py both split > gpurun_out/layers_$o.txt 2>&1; grep -E "^==|total conv" gpurun_out/layers_$o.txt; done
for o in fixed alternate; do FT_TILE_ORDER=$o timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$o', d['value'], d['e2e']['value'], d['ms_per_step'])"; done
