"""Inference entry point of the reference's tools/flownet/main.py (`--inference [--save_flow]`, :53-57,118-123 -> inference(),
:340-390) on the B200 path.  Training, validation losses, DataParallel and the Sintel / Chairs / Things dataset classes
are out of scope; kept: the flag names, the `ImagesFromFolder` inference dataset (consecutive frame
pairs of a folder, centre-cropped to a multiple of 64; lib/flownet/model/datasets.py:320-365), the checkpoint format
({'state_dict', 'epoch', 'best_EPE'}, utils/tools.py:138-143) and the output layout
`<save>/inference/<name>.epoch-<e>-flow-field/%06d.flo` (Middlebury .flo, byte-identical to the reference writer).
`--save_flow_image` additionally writes the Middlebury colour coding (flowlib.flow_to_image, computed on the GPU) as PNG.

    python tools/flownet/main.py --inference --save_flow --model FlowNet2S --resume FlowNet2-S_checkpoint.pth.tar \
        --inference_dataset ImagesFromFolder --inference_dataset_root frames/ --inference_dataset_iext png
"""
import argparse
import os
import sys
from glob import glob
from os.path import join

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from flowtrack.pytorch_b200.flownet.model import models  # noqa: E402
from flowtrack.pytorch_b200.flownet.utils import flow_utils  # noqa: E402

MODELS = ["FlowNet2", "FlowNet2C", "FlowNet2S", "FlowNet2SD", "FlowNet2CS", "FlowNet2CSS"]


def read_gen(file_name):
    """lib/flownet/utils/frame_utils.py:6-18 (scipy.misc.imread is gone; cv2 reads the same RGB bytes)."""
    ext = os.path.splitext(file_name)[-1]
    if ext in (".png", ".jpeg", ".ppm", ".jpg"):
        import cv2
        im = cv2.imread(file_name, cv2.IMREAD_UNCHANGED)
        if im.ndim == 2:
            im = np.repeat(im[:, :, None], 3, axis=2)
        return im[:, :, :3][:, :, ::-1]
    if ext in (".bin", ".raw"):
        return np.load(file_name)
    if ext == ".flo":
        return flow_utils.readFlow(file_name).astype(np.float32)
    return []


class StaticCenterCrop(object):
    """datasets.py:23-28."""

    def __init__(self, image_size, crop_size):
        self.th, self.tw = crop_size
        self.h, self.w = image_size

    def __call__(self, img):
        return img[(self.h - self.th) // 2:(self.h + self.th) // 2, (self.w - self.tw) // 2:(self.w + self.tw) // 2, :]


class ImagesFromFolder(torch.utils.data.Dataset):
    """datasets.py:320-365, inference branch (is_cropped=False)."""

    def __init__(self, args, is_cropped=False, root="/path/to/frames/only/folder", iext="png", replicates=1):
        self.args, self.replicates = args, replicates
        self.render_size = list(args.inference_size)
        images = sorted(glob(join(root, "*." + iext)))
        self.image_list = [[images[i], images[i + 1]] for i in range(len(images) - 1)]
        self.size = len(self.image_list)
        if self.size == 0:
            raise SystemExit("ImagesFromFolder: need at least two *.%s frames under %s" % (iext, root))
        self.frame_size = read_gen(self.image_list[0][0]).shape
        if self.render_size[0] < 0 or self.render_size[1] < 0 or self.frame_size[0] % 64 or self.frame_size[1] % 64:
            self.render_size[0] = (self.frame_size[0] // 64) * 64
            self.render_size[1] = (self.frame_size[1] // 64) * 64
        args.inference_size = self.render_size

    def __getitem__(self, index):
        index = index % self.size
        img1, img2 = read_gen(self.image_list[index][0]), read_gen(self.image_list[index][1])
        cropper = StaticCenterCrop(img1.shape[:2], self.render_size)
        images = np.array([cropper(img1), cropper(img2)]).transpose(3, 0, 1, 2)
        images = torch.from_numpy(images.astype(np.float32))
        return [images], [torch.zeros(images.size()[0:1] + (2,) + images.size()[-2:])]

    def __len__(self):
        return self.size * self.replicates


DATASETS = {"ImagesFromFolder": ImagesFromFolder}


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument("--rgb_max", type=float, default=255.)
    parser.add_argument("--number_workers", "-nw", "--num_workers", type=int, default=0)
    parser.add_argument("--number_gpus", "-ng", type=int, default=1, help="number of GPUs to use (frame pairs are sharded; 1 here)")
    parser.add_argument("--seed", type=int, default=1)
    parser.add_argument("--name", default="run", type=str, help="a name to append to the save directory")
    parser.add_argument("--save", "-s", default="./work", type=str, help="directory for saving")
    parser.add_argument("--inference", action="store_true")
    parser.add_argument("--inference_size", type=int, nargs="+", default=[-1, -1], help="spatial size divisible by 64; (-1,-1): largest valid size")
    parser.add_argument("--inference_batch_size", type=int, default=1)
    parser.add_argument("--inference_n_batches", type=int, default=-1)
    parser.add_argument("--save_flow", action="store_true", help="save predicted flows to file")
    parser.add_argument("--save_flow_image", action="store_true", help="also save the Middlebury colour coding of every flow as PNG")
    parser.add_argument("--resume", default="", type=str, metavar="PATH", help="path to latest checkpoint (default: none)")
    parser.add_argument("--fp16", action="store_true", help="single-pass fp16 operands (FT_PREC_FAST) instead of the fp32-grade split")
    parser.add_argument("--model", default="FlowNet2", choices=MODELS)
    parser.add_argument("--model_batchNorm", action="store_true")
    parser.add_argument("--model_div_flow", type=float, default=20.)
    parser.add_argument("--inference_dataset", default="ImagesFromFolder", choices=sorted(DATASETS))
    parser.add_argument("--inference_dataset_root", default="./frames", type=str)
    parser.add_argument("--inference_dataset_iext", default="png", type=str)
    parser.add_argument("--inference_dataset_replicates", default=1, type=int)
    return parser


def inference(args, epoch, data_loader, model, offset=0):
    """tools/flownet/main.py:340-390: run the model over the loader; with --save_flow write one .flo per pair."""
    model.eval()
    flow_folder = "{}/{}.epoch-{}-flow-field".format(args.inference_dir, args.name.replace("/", "."), epoch)
    if args.save_flow or args.save_flow_image:
        os.makedirs(flow_folder, exist_ok=True)
    n_batches = np.inf if args.inference_n_batches < 0 else args.inference_n_batches
    statistics = []
    for batch_idx, (data, target) in enumerate(data_loader):
        with torch.no_grad():
            output = model(data[0].cuda(non_blocking=True))                                  # [B,2,H,W]
        # the reference reports the L1 / L2 norms of the flow against all-zero targets when no ground truth exists (:362-365)
        statistics.append([output.abs().mean().item(), output.pow(2).sum(1).sqrt().mean().item()])
        if args.save_flow:
            for i in range(output.shape[0]):
                flow_utils.writeFlow(join(flow_folder, "%06d.flo" % (batch_idx * args.inference_batch_size + i)), output[i])
        if args.save_flow_image:
            import cv2
            imgs = flow_utils.flow_to_image_cuda(output).cpu().numpy()
            for i in range(output.shape[0]):
                cv2.imwrite(join(flow_folder, "%06d.png" % (batch_idx * args.inference_batch_size + i)), imgs[i][:, :, ::-1])
        if batch_idx == n_batches - 1:
            break
    mean = np.array(statistics).mean(axis=0) if statistics else np.zeros(2)
    print("Inference Averages for Epoch {}: L1: {:.3f}, EPE: {:.3f}".format(epoch, mean[0], mean[1]))
    return flow_folder


def main(argv=None):
    args = build_parser().parse_args(argv)
    if not args.inference:
        raise SystemExit("only --inference is implemented on the B200 path (training / validation are out of scope)")
    args.grads = {}
    args.inference_dir = "{}/inference".format(args.save)
    args.crop_size = args.inference_size
    dataset = DATASETS[args.inference_dataset](args, False, root=args.inference_dataset_root, iext=args.inference_dataset_iext,
                                               replicates=args.inference_dataset_replicates)
    print("Inference Dataset: {}  Input: {}".format(args.inference_dataset, [list(x.size()) for x in dataset[0][0]]))
    loader = torch.utils.data.DataLoader(dataset, batch_size=args.inference_batch_size, shuffle=False, num_workers=args.number_workers, pin_memory=True)
    model = getattr(models, args.model)(args, batchNorm=args.model_batchNorm, div_flow=args.model_div_flow)
    print("Number of parameters: {}".format(sum(p.numel() for p in model.parameters())))
    epoch = 1
    if args.resume and os.path.isfile(args.resume):
        checkpoint = torch.load(args.resume, map_location="cpu")
        model.load_state_dict(checkpoint["state_dict"])
        epoch = checkpoint.get("epoch", 1)
        print("Loaded checkpoint '{}' (at epoch {})".format(args.resume, epoch))
    elif args.resume:
        raise SystemExit("No checkpoint found at '{}'".format(args.resume))
    else:
        print("Random initialization")
    model = model.cuda().eval()
    if args.fp16:
        model.precision = "fast"
    folder = inference(args, epoch - 1 if args.resume else 0, loader, model)
    print("flows under", folder)
    return 0


if __name__ == "__main__":
    sys.exit(main())
