#!/bin/bash
# sanitizer pass over the small SIMT kernels of ft_track.cu / ft_corr.cu's re-layout kernels (the tcgen05 kernels are not instrumentable in reasonable time)
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_track.py -m gpu -x -q -k "nms or crops or flip_test_forward_matches" -p no:cacheprovider > gpurun_out/sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitizer_memcheck.txt
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_track.py -m gpu -x -q -k "nms_vs_oracle or nms_oks" -p no:cacheprovider > gpurun_out/sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/sanitizer_racecheck.txt
