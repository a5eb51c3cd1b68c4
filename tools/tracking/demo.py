"""Pose estimation in video: the per-frame step of the reference's tools/tracking/demo.py (process_frame, :35-42) on the
B200 path.  The reference script is a specification by intent (it does not parse: `type=str, ,` :24, `paeser` :30); this
entry point keeps its argument names and the process_frame signature.  Detection (Faster R-CNN) is out of scope: pass a
callable with the im_detect contract as `detect_net`, or None to track the propagated boxes only.

    python tools/tracking/demo.py --pose_backbone 50 --pose_model pose.pth --flow_net FlowNet2S --flow_model flow.pth \
        --frames f0.jpg f1.jpg ... [--boxes x1,y1,x2,y2 ...]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from flowtrack.pytorch_b200.flownet.model import models as flow_models  # noqa: E402
from flowtrack.pytorch_b200.pose import models as pose_models  # noqa: E402
from flowtrack.pytorch_b200.tracking.flow_utils import box_propagation  # noqa: E402
from flowtrack.pytorch_b200.tracking.net_utils import detect, flow_est, pose_est  # noqa: E402

PERSON_ID = 1


def parse_args(argv=None):
    parser = argparse.ArgumentParser(description="Pose Estimation in Video")
    parser.add_argument("--detect_net", default="res101", type=str, help="(detector is out of scope; kept for CLI compatibility)")
    parser.add_argument("--detect_dataset", default="coco", type=str, help="coco, pascal_voc")
    parser.add_argument("--detect_model", type=str, help="human detection model (unused)")
    parser.add_argument("--pose_backbone", default=152, type=int, help="50, 101, 152")
    parser.add_argument("--pose_model", type=str, help="pose estimation checkpoint ({'state_dict': ...})")
    parser.add_argument("--flow_net", type=str, default="FlowNet2S", help="FlowNet2, FlowNet2C, FlowNet2S, FlowNet2SD, FlowNet2CS, FlowNet2CSS")
    parser.add_argument("--flow_model", type=str, help="optical flow checkpoint ({'state_dict': ...})")
    parser.add_argument("--frames", nargs="*", default=[], help="image files, in temporal order")
    parser.add_argument("--boxes", nargs="*", default=[], help="person boxes of the first frame: x1,y1,x2,y2[,score]")
    return parser.parse_args(argv)


def process_frame(detect_net, pose_net, flownet, prev_im, cur_im, prev_dets, prev_keypoints):
    """tools/tracking/demo.py:35-42, verbatim data flow."""
    flow = flow_est(flownet, prev_im, cur_im)
    prop_boxes = box_propagation(prev_keypoints, flow)
    prop_dets = np.concatenate((prop_boxes, prev_dets[:, 4][:, np.newaxis]), axis=1)
    cur_dets = detect(detect_net, cur_im, PERSON_ID, prop_dets=prop_dets)
    cur_keypoints = pose_est(pose_net, cur_im, cur_dets[:, :4])
    return cur_dets, cur_keypoints


def build_nets(args):
    pose_net = pose_models.deconv("resnet%d" % args.pose_backbone, 17, False)
    if args.pose_model:
        pose_net.load_state_dict(torch.load(args.pose_model, map_location="cpu")["state_dict"])
    flow_net = getattr(flow_models, args.flow_net)(None)
    if args.flow_model:
        flow_net.load_state_dict(torch.load(args.flow_model, map_location="cpu")["state_dict"])
    return pose_net.eval().cuda(), flow_net.eval().cuda()


if __name__ == "__main__":
    import cv2
    args = parse_args()
    PERSON_ID = 15 if args.detect_dataset == "pascal_voc" else 1
    pose_net, flow_net = build_nets(args)
    frames = [cv2.imread(f) for f in args.frames]
    dets = np.array([[float(v) for v in b.split(",")] + ([1.0] if b.count(",") == 3 else []) for b in args.boxes], np.float32).reshape(-1, 5)
    kps = pose_est(pose_net, frames[0], dets[:, :4]) if len(frames) else None
    for i in range(1, len(frames)):
        dets, kps = process_frame(None, pose_net, flow_net, frames[i - 1], frames[i], dets, kps)
        print("frame %d: %d people" % (i, dets.shape[0]))
        for d, k in zip(dets, kps):
            print("  box %s score %.3f  keypoints %s" % (np.round(d[:4], 1).tolist(), d[4], np.round(k[:, :2], 1).tolist()))
