"""BASELINE config 5 on one GPU (dev tool): PoseResNet-152 on P=4 crops 384x288 per pair + full FlowNet2 at 384x512.
Usage: config5_bench.py [pairs_per_gpu ...]   (config 5 = 8 pairs over 8 GPUs = 1 pair per GPU; also timed: 8 pairs on one GPU)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from flowtrack.pytorch_b200.flownet.model import models as flow_models  # noqa: E402
from flowtrack.pytorch_b200.pose import models as pose_models  # noqa: E402
from flowtrack.pytorch_b200.tracking import pipeline as pl  # noqa: E402
from flowtrack.pytorch_b200.tracking import synthetic  # noqa: E402

GFLOP_PER_PAIR = 198.24 + 0.694 + 4 * 57.147
torch.manual_seed(0)
pose = synthetic.condition_pose_weights(pose_models.deconv("resnet152", 17, False), 0).cuda()
flow = flow_models.FlowNet2(None).cuda()
for pairs in [int(a) for a in sys.argv[1:]] or [1, 8]:
    pipe = pl.FramePairPipeline(pose, flow, pairs, 4, (384, 512), (384, 288), "cuda:0", crops_from_frames=True)
    d = synthetic.frame_pair_batch(pairs, 4, (384, 512), (384, 288), seed=1)
    pipe.process(d["pair"], None, d["centers"], d["scales"], d["prev_keypoints"])
    for _ in range(3):
        pipe.step_device()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 20
    e0.record()
    for _ in range(steps):
        pipe.step_device()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print("config5 (R152 384x288 x4 crops + FlowNet2 384x512), %d pair(s)/GPU: %.3f ms/step, %.1f frame-pairs/s per GPU, %.1f TFLOP/s algorithmic, %d launches/step"
          % (pairs, ms, pairs / ms * 1e3, pairs * GFLOP_PER_PAIR / ms, pipe.launches_per_step))
    del pipe
    torch.cuda.empty_cache()
