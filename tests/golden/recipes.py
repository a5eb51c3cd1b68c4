"""Seeded weight/input recipes shared by make_golden.py (reference side) and the tests (oracle side)."""
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
