#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_track.py -m gpu -x -q 2>&1 | tail -15
