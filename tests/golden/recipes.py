"""Seeded weight/input recipes shared by make_golden.py (reference side) and the tests (oracle side)."""
import numpy as np
import torch

from oracle import flow_oracle, pose_oracle

POSE_CASES = {50: ((128, 96), 2), 101: ((64, 64), 1), 152: ((64, 64), 1)}
FLOW_NAMES = ("FlowNet2S", "FlowNet2C", "FlowNet2SD", "FlowNet2CS", "FlowNet2CSS", "FlowNet2")


def pose_oracle_model(depth, calib_hw=None, calib_batch=4):
    hw, _ = POSE_CASES.get(depth, ((64, 64), 1))
    torch.manual_seed(1000 + depth)
    m = pose_oracle.deconv("resnet%d" % depth, 17, False)
    return pose_oracle.condition_weights(m, seed=depth, calib_hw=calib_hw or hw, calib_batch=calib_batch)


def pose_input(depth):
    hw, nb = POSE_CASES[depth]
    return torch.randn(nb, 3, hw[0], hw[1], generator=torch.Generator().manual_seed(100 + depth))


def flow_oracle_model(name):
    torch.manual_seed(7)
    return getattr(flow_oracle, name)(None).eval()


def flow_input(batch=1, hw=(64, 128), seed=11):
    return torch.rand(batch, 3, 2, hw[0], hw[1], generator=torch.Generator().manual_seed(seed)) * 255.0


def decode_heatmaps():
    g = torch.Generator().manual_seed(5)
    hm = torch.randn(3, 17, 64, 48, generator=g)
    hm[0, 3] = -hm[0, 3].abs()      # all-negative map: score <= 0 -> coords zeroed
    hm[1, 5, 0, 7] = 9.0            # maximum on the border: no quarter-pixel shift
    return hm


# ---- tracking glue (person crops) ----
TRACK_RES = (64, 48)
TRACK_BOXES = np.array([[10.0, 5.0, 60.0, 90.0], [50.5, 20.25, 120.0, 80.0], [-20.0, -10.0, 40.0, 50.5], [100.0, 60.0, 127.0, 95.0]])


def track_frame():
    """A deterministic 96x128 BGR uint8 frame (smooth gradients + noise, so interpolation errors are visible)."""
    rng = np.random.default_rng(7)
    yy, xx = np.mgrid[0:96, 0:128]
    base = np.stack([(xx * 2) % 256, (yy * 3) % 256, (xx + yy) % 256], axis=2)
    return ((base + rng.integers(0, 64, base.shape)) % 256).astype(np.uint8)


def flip_heatmaps():
    return np.random.default_rng(9).standard_normal((2, 17, 8, 6)).astype(np.float32)


COCO_DELTA = 2 * np.array([.26, .25, .25, .35, .35, .79, .79, .72, .72, .62, .62, 1.07, 1.07, .87, .87, .89, .89]) / 10.0


def oks_poses(n=40, seed=12):
    """n synthetic 17-joint poses in clusters (near-duplicates around a few people) with distinct pose scores."""
    rng = np.random.default_rng(seed)
    centres = rng.uniform(50, 400, (6, 1, 2))
    base = centres + rng.uniform(-40, 40, (6, 17, 2))
    preds = []
    for i in range(n):
        c = i % 6
        xy = base[c] + rng.normal(0, 1.5 + 3.0 * (i // 6), (17, 2))
        sc = rng.uniform(0.05, 1.0, (17, 1))
        preds.append({"score": float(1.0 - i * 0.011 + 0.003 * rng.random()), "joints": np.concatenate([xy, sc], 1).astype(np.float32),
                      "area": float(rng.uniform(3000, 9000))})
    return preds


def nms_boxes(n, seed):
    """n seeded person-like boxes with distinct scores (clusters of overlapping boxes)."""
    rng = np.random.default_rng(seed)
    centres = rng.uniform(40, 600, (max(1, n // 6), 2))
    c = centres[rng.integers(0, len(centres), n)] + rng.normal(0, 12, (n, 2))
    wh = rng.uniform(30, 160, (n, 2))
    sc = (rng.permutation(n).astype(np.float32) + 1) / (n + 1)
    return np.concatenate([c - wh / 2, c + wh / 2, sc[:, None]], axis=1).astype(np.float32)


NMS_CASES = [(1, 0, 0.5), (17, 1, 0.3), (120, 2, 0.5), (300, 3, 0.7), (1000, 4, 0.3)]


def flowvis_flows():
    """Seeded [H,W,2] float32 flows for the colour-coding fixtures: smooth rotation field + noise (every wheel sector),
    a small-magnitude field, and one with unknown (> 1e7), zero and NaN-free extreme entries."""
    rng = np.random.default_rng(21)
    yy, xx = np.mgrid[0:48, 0:64].astype(np.float32)
    rot = np.stack([(yy - 24) * 0.7, -(xx - 32) * 0.5], axis=2) + rng.normal(0, 0.3, (48, 64, 2)).astype(np.float32)
    small = rng.normal(0, 0.01, (16, 24, 2)).astype(np.float32)
    mixed = rng.normal(0, 12.0, (32, 40, 2)).astype(np.float32)
    mixed[3, 4] = (2e7, 1.0)
    mixed[5, 6] = (0.0, -3e8)
    mixed[8:12, 10:14] = 0.0
    return [rot.astype(np.float32), small, mixed]
