#!/bin/bash
# 8-GPU weak-scaling check of bench.py (torchrun, NCCL)
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -3 gpurun_out/bench_n$N.err; cut -c1-330 gpurun_out/bench_n$N.json
