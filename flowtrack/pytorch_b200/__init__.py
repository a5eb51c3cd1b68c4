"""flowtrack.pytorch_b200 -- B200 (sm_100a) implementation of FlowTrack's per-frame-pair inference
hot path behind the reference's own module interfaces:

    from flowtrack.pytorch_b200.pose import models            # models.deconv('resnet50', 17, False)
    from flowtrack.pytorch_b200.flownet.model import models   # models.FlowNet2S(args), FlowNet2(args), ...
    from flowtrack.pytorch_b200.tracking import net_utils, flow_utils

All arithmetic runs in libflowtrack_b200.so (hand-written CUDA, C ABI in include/flowtrack_b200.h).
There is no CPU / PyTorch fallback: forward() on a non-CUDA tensor, or without the built library, raises.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
