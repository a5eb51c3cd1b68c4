#!/bin/bash
# round-2 check job: gpu tests + smoke + short bench (logs under gpurun_out/)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x "$@" > gpurun_out/r02_pytest_gpu.txt 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke.txt
