"""Run PoseResNet forward(s) without CUDA graphs so ncu sees one launch per layer (dev tool).
Usage: pose_once.py [depth] [batch] [precision] [n_forwards]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from flowtrack.pytorch_b200 import _lib  # noqa: E402
from flowtrack.pytorch_b200.pose import models  # noqa: E402

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 50
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
prec = sys.argv[3] if len(sys.argv) > 3 else "split"
n = int(sys.argv[4]) if len(sys.argv) > 4 else 2
torch.manual_seed(0)
m = models.deconv("resnet%d" % depth, 17, False)
with torch.no_grad():
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_var.fill_(1.0)
            mod.weight.uniform_(0.2, 0.6)
m.cuda()
m.precision = prec
m.use_cuda_graph = False
x = torch.randn(B, 3, 256, 192, device="cuda")
for i in range(n):
    _lib.reset_launch_count()
    if i == n - 1:
        torch.cuda.profiler.start()      # `ncu --profile-from-start off` captures only this last forward (not the warm-up / autotune launches)
    m(x)
    torch.cuda.synchronize()
    if i == n - 1:
        torch.cuda.profiler.stop()
print("launches per forward:", _lib.launch_count())
plan = list(m._plans.values())[0][0]
for i, c in enumerate(plan.calls):
    if c[0] == "conv2d":
        d = c[3]
        kind = "deconv" if d.kind else "conv"
        ho = d.out.h if d.out.hi else 0
        print("%3d %-6s %dx%d s%d cin_p=%d cout=%d in=%dx%d BN=%d k_pad=%d" % (i, kind, d.kh, d.kw, d.stride, d.in_.c, d.cout, d.in_.h, d.in_.w, d.block_n, d.k_pad))
    else:
        print("%3d %s" % (i, c[0]))
