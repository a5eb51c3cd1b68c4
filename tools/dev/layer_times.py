"""Per-layer device time of the pose / flow plans (dev tool): CUDA-event pair around every C-ABI call, un-graphed.
Usage: layer_times.py [pose|flow|both] [precision] [pose_batch] [flow_batch]  -> table on stdout"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from flowtrack.pytorch_b200 import _lib  # noqa: E402
from flowtrack.pytorch_b200.flownet.model import models as flow_models  # noqa: E402
from flowtrack.pytorch_b200.pose import models as pose_models  # noqa: E402
from flowtrack.pytorch_b200.tracking import synthetic  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "both"
prec = sys.argv[2] if len(sys.argv) > 2 else "split"
PB = int(sys.argv[3]) if len(sys.argv) > 3 else 64
FB = int(sys.argv[4]) if len(sys.argv) > 4 else 16
PEAK = 1630.9    # measured burst bf16 TFLOP/s (MEASURED_PEAKS.json)


def table(name, plan):
    for _ in range(2):
        plan.enqueue()
    torch.cuda.synchronize()
    per = plan.time_calls_graphed(reps=5) if os.environ.get("LT_GRAPHED", "1") != "0" else plan.time_calls(reps=5)
    tot = sum(per)
    tot_fl = 0.0
    print("== %s (%s): %d calls, %.3f ms summed" % (name, prec, len(per), tot))
    for i, (what, _fn, _a, keep) in enumerate(plan.calls):
        if _fn is None:
            continue
        if what == "conv2d":
            d = keep
            fl = plan.flops[i]
            tot_fl += fl
            M = d.in_.n * (d.in_.h * d.in_.w if d.kind else ((d.in_.h + 2 * d.pad - d.kh) // d.stride + 1) * ((d.in_.w + 2 * d.pad - d.kw) // d.stride + 1))
            tiles = (4 if d.kind else 1) * ((M + 127) // 128) * (d.cout_pad // d.block_n)
            print("%3d %-6s k%d s%d cin_p=%-4d cout=%-4d in=%3dx%-3d M=%-7d BN=%-3d%-4s kpad=%-5d tiles=%-5d %8.1f us %7.1f TF/s alg (%4.1f%% of peak)"
                  % (i, "deconv" if d.kind else "conv", d.kh, d.stride, d.in_.c, d.cout, d.in_.h, d.in_.w, M, d.block_n, ("Q" if d.cta_pair == 2 else "P" if d.cta_pair else " ") + ("2" if d.epilogue_buffers == 2 else " ") + ("k%d" % d.splitk if d.splitk > 1 else "") + ("o2" if d.ctas_per_sm == 2 else ""), d.k_pad, tiles,
                     per[i] * 1e3, fl / per[i] / 1e9, fl / per[i] / 1e9 / PEAK * 100))
        else:
            print("%3d %-40s %8.1f us" % (i, what, per[i] * 1e3))
    print("   total conv algorithmic %.1f GFLOP -> %.1f TF/s over the summed time" % (tot_fl / 1e9, tot_fl / tot / 1e9))


torch.manual_seed(0)
if which in ("pose", "both"):
    m = synthetic.condition_pose_weights(pose_models.deconv("resnet50", 17, False), 0).cuda()
    m.precision = prec
    plan, _x, _o = m.plan_for(PB, 256, 192, torch.device("cuda"))
    table("PoseResNet-50 B=%d 256x192" % PB, plan)
if which in ("flow", "both"):
    f = flow_models.FlowNet2S(None).cuda()
    f.precision = prec
    plan, _x, _o = f.plan_for(FB, 384, 512, torch.device("cuda"))
    table("FlowNet2S B=%d 384x512" % FB, plan)
