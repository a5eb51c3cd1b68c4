"""Correlation kernel timings at BASELINE config 3 (dev tool): prints bench.bench_correlation()'s record."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
from flowtrack.pytorch_b200 import _lib  # noqa: E402

print(json.dumps(bench.bench_correlation(torch, _lib, bench.load_peaks()), indent=1))
