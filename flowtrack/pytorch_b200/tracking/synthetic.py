"""Seeded synthetic weights and inputs for the joint pipeline (bench.py, __graft_entry__.smoke, tests).

There are no datasets or checkpoints offline, so the workload is synthetic: random-init networks of the
reference architectures and random frames / crops of the BASELINE shapes.  The reference's default
pose init makes the heat-maps ~1e-4 in magnitude (deconv / head weights are N(0, 1e-3), pose_deconv.py:53,62), so
`condition_pose_weights` rescales the head the way SURVEY.md 8(d) describes and gives every BatchNorm non-trivial
running statistics, in closed form (no training-mode forward: this package has no training path)."""
import math

import numpy as np
import torch


@torch.no_grad()
def condition_pose_weights(model, seed=0):
    g = torch.Generator().manual_seed(seed)
    for name, m in model.named_modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            lo, hi = (0.1, 0.3) if name.endswith("bn3") else (0.5, 1.5)
            m.weight.copy_(torch.rand(m.weight.shape, generator=g) * (hi - lo) + lo)
            m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
            m.running_mean.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.bias.shape, generator=g) * 0.5 + 0.75)
        elif isinstance(m, torch.nn.ConvTranspose2d):
            m.weight.copy_(torch.randn(m.weight.shape, generator=g) * math.sqrt(2.0 / (m.in_channels * 4)))
    model.heatmap.weight.copy_(torch.randn(model.heatmap.weight.shape, generator=g) / 16.0)
    model.heatmap.bias.copy_(torch.randn(model.heatmap.bias.shape, generator=g) * 0.1)
    if hasattr(model, "refresh"):
        model.refresh()
    return model


def frame_pair_batch(pairs, people, frame_hw=(384, 512), crop_hw=(256, 192), num_joints=17, seed=0):
    """One batch of the joint workload as CPU fp32 tensors:
    pair [pairs,3,2,H,W] U(0,255); crops [pairs*people,3,h,w] N(0,1); person boxes on a fixed grid (96x192 px) ->
    centers/scales by the pose_est rule (net_utils.py:49-50); previous-frame key-points inside the boxes with
    N(0.5,0.5) scores (about one joint in six is below the score threshold and is skipped by the propagation)."""
    g = torch.Generator().manual_seed(seed)
    H, W = frame_hw
    n = pairs * people
    pair = torch.rand(pairs, 3, 2, H, W, generator=g) * 255.0
    crops = torch.randn(n, 3, crop_hw[0], crop_hw[1], generator=g)
    boxes = np.zeros((n, 4), np.float32)
    bw, bh = min(96.0, W / 2), min(192.0, H / 2)
    for i in range(n):
        k = i % people
        x0 = (W - bw) * (k + 0.5) / people
        y0 = (H - bh) * (0.25 + 0.5 * ((k * 7) % 3) / 2.0)
        boxes[i] = (x0, y0, x0 + bw, y0 + bh)
    centers = np.stack((boxes[:, [0, 2]].mean(1), boxes[:, [1, 3]].mean(1)), axis=1).astype(np.float32)
    scales = np.maximum(boxes[:, 3] - boxes[:, 1], (boxes[:, 2] - boxes[:, 0]) / crop_hw[1] * crop_hw[0]).astype(np.float32)
    u = torch.rand(n, num_joints, 2, generator=g)
    kx = torch.from_numpy(boxes[:, None, 0]) + u[:, :, 0] * bw
    ky = torch.from_numpy(boxes[:, None, 1]) + u[:, :, 1] * bh
    score = torch.randn(n, num_joints, generator=g) * 0.5 + 0.5
    prev = torch.stack((kx, ky, score), dim=2).float().contiguous()
    return dict(pair=pair, crops=crops, centers=torch.from_numpy(centers), scales=torch.from_numpy(scales), prev_keypoints=prev,
                boxes=torch.from_numpy(boxes))
