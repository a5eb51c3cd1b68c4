"""Optical flow of one image pair: the reference's tools/flownet/demo.py on the B200 path -- same arguments, writes
<save>/output.flo (Middlebury format, lib/flownet/utils/flow_utils.py).  The colour-wheel PNG of the reference needs
matplotlib (absent here) and is skipped.

    python tools/flownet/demo.py --input1 img0.ppm --input2 img1.ppm --resume FlowNet2-S_checkpoint.pth.tar --model FlowNet2S
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from flowtrack.pytorch_b200.flownet.model import models  # noqa: E402
from flowtrack.pytorch_b200.flownet.utils import flow_utils  # noqa: E402


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--input1", "-i", default="samples/img0.ppm", type=str, help="first input image")
    parser.add_argument("--input2", "-p", default="samples/img1.ppm", type=str, help="second input image")
    parser.add_argument("--save", "-s", default="results", type=str, help="directory for saving")
    parser.add_argument("--resume", default="", type=str, metavar="PATH", help="path to checkpoint ({'state_dict','epoch','best_EPE'})")
    parser.add_argument("--number_gpus", "-ng", type=int, default=1)
    parser.add_argument("--fp16", action="store_true", help="single-pass fp16 operands (FT_PREC_FAST) instead of the fp32-grade split")
    parser.add_argument("--rgb_max", type=float, default=255.)
    parser.add_argument("--model", default="FlowNet2S", choices=["FlowNet2", "FlowNet2C", "FlowNet2S", "FlowNet2SD", "FlowNet2CS", "FlowNet2CSS"])
    args = parser.parse_args(argv)
    args.grads = {}
    model = getattr(models, args.model)(args)
    if args.resume and os.path.isfile(args.resume):
        checkpoint = torch.load(args.resume, map_location="cpu")
        model.load_state_dict(checkpoint["state_dict"])
        print("Loaded checkpoint '%s' (at epoch %s)" % (args.resume, checkpoint.get("epoch")))
    else:
        print("No checkpoint found at '%s'" % args.resume)
        return 1
    os.makedirs(args.save, exist_ok=True)
    model = model.cuda().eval()
    if args.fp16:
        model.precision = "fast"
    import cv2
    im1, im2 = (cv2.imread(f)[:, :, ::-1] for f in (args.input1, args.input2))          # H x W x 3 (RGB)
    ims = torch.from_numpy(np.array([[im1, im2]]).transpose((0, 4, 1, 2, 3)).astype(np.float32)).cuda()   # B x 3 x 2 x H x W
    pred_flow = model(ims)[0]                                                           # 2 x H x W
    flow_utils.writeFlow(os.path.join(args.save, "output.flo"), pred_flow)
    print("wrote", os.path.join(args.save, "output.flo"), tuple(pred_flow.shape))
    return 0


if __name__ == "__main__":
    sys.exit(main())
