import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# the per-layer tile autotuner (engine.Plan.autotune) costs ~1 s per plan; the suite builds dozens of small plans, so it is
# off by default here and exercised by tests/test_gpu_conv.py::test_autotune_* explicitly
os.environ.setdefault("FT_AUTOTUNE", "0")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a box without CUDA (or without the built library) skips the gpu-marked tests instead of failing
    inside torch.  On a GPU box nothing is skipped here: a missing libflowtrack_b200.so must fail loudly there."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="needs CUDA (sm_100a) + the built libflowtrack_b200.so")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_outputs.npz"))
