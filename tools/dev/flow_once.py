"""Run FlowNet2S forward(s) without CUDA graphs so ncu sees one launch per layer (dev tool).
Usage: flow_once.py [batch] [precision] [n_forwards]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from flowtrack.pytorch_b200 import _lib  # noqa: E402
from flowtrack.pytorch_b200.flownet.model import models  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
prec = sys.argv[2] if len(sys.argv) > 2 else "split"
n = int(sys.argv[3]) if len(sys.argv) > 3 else 2
torch.manual_seed(0)
m = models.FlowNet2S(None).cuda()
m.precision = prec
m.use_cuda_graph = False
x = torch.rand(B, 3, 2, 384, 512, device="cuda") * 255
for i in range(n):
    _lib.reset_launch_count()
    if i == n - 1:
        torch.cuda.profiler.start()      # `ncu --profile-from-start off` captures only this last forward (not the warm-up / autotune launches)
    m(x)
    torch.cuda.synchronize()
    if i == n - 1:
        torch.cuda.profiler.stop()
print("launches per forward:", _lib.launch_count())
