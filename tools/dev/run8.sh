#!/bin/bash
mkdir -p gpurun_out
python tools/dev/acc_probe.py > gpurun_out/acc_probe.log 2>&1; cat gpurun_out/acc_probe.log
