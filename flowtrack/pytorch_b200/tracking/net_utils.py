"""pose_est / flow_est with the reference signatures (lib/tracking/net_utils.py:36-92).  `detect` (Faster R-CNN)
is out of scope: person boxes are an input of the hot path.

Fixes relative to the literal source, each documented in SURVEY.md row A11: the batching loop's
`end = max(...)` (net_utils.py:61) is `min`; the imports (`pose.transforms`, misspelt `transfrom_image`, :11-12)
resolve to this package; `Variable(volatile=True)` is torch.no_grad().  Kept as in the reference: no mean/std
normalisation of the crops, scale = max(box_h, box_w/192*256) without margin, cv2.warpAffine (bilinear, zero
border) for the crops, decode = final_preds(adjust_coords=True)."""
import numpy as np
import torch

from .. import _lib, engine


def get_transform(center, scale, res):
    """lib/pose/utils/transforms.py:173-184 (factor=1, rot=0, delta=0 branch)."""
    t = np.eye(3)
    t[0, 0] = t[1, 1] = res[0] / scale
    t[0, 2] = -res[0] * center[0] / scale + 0.5 * res[1]
    t[1, 2] = -res[0] * center[1] / scale + 0.5 * res[0]
    return t


def transform_image(img, center, scale, res):
    """lib/pose/utils/transforms.py:231-240: affine crop of an HxWxC image to res=(h,w); returns a CxHxW tensor."""
    import cv2
    if torch.is_tensor(img):
        img = img.numpy()
        if img.ndim > 2:
            img = img.transpose((1, 2, 0))
    out = cv2.warpAffine(img, get_transform(center, scale, res)[:2], (res[1], res[0]))
    if img.ndim > 2:
        out = out.transpose((2, 0, 1))
    return torch.from_numpy(np.ascontiguousarray(out))


def boxes_to_center_scale(boxes, inp_res=(256, 192)):
    """net_utils.py:49-50."""
    boxes = np.asarray(boxes, dtype=np.float64)
    centers = np.stack((boxes[:, [0, 2]].mean(1), boxes[:, [1, 3]].mean(1)), axis=1)
    scales = np.maximum(boxes[:, 3] - boxes[:, 1], (boxes[:, 2] - boxes[:, 0]) / inp_res[1] * inp_res[0])
    return centers, scales


def final_preds_cuda(heatmaps, centers, scales, adjust_coords=True, out=None):
    """max_preds + final_preds + transform_preds (lib/pose/utils/evaluation.py:11-35) in one kernel.
    heatmaps: CUDA fp32 [N,C,h,w]; centers [N,2], scales [N] CUDA fp32 -> CUDA fp32 [N,C,3] = (x, y, score)."""
    assert heatmaps.is_cuda and heatmaps.dtype == torch.float32 and heatmaps.is_contiguous()
    N, C, h, w = heatmaps.shape
    if out is None:
        out = torch.empty(N, C, 3, device=heatmaps.device, dtype=torch.float32)
    _lib.check(_lib.lib().ft_decode_keypoints(heatmaps.data_ptr(), N, C, h, w, centers.data_ptr(), scales.data_ptr(), int(bool(adjust_coords)),
                                              out.data_ptr(), engine.stream_ptr()), "decode_keypoints")
    return out


@torch.no_grad()
def pose_est(net, im, boxes, inp_res=(256, 192), max_batch=32, flip_test=False):
    """im: HxWx3 ndarray (BGR, as loaded by cv2); boxes: ndarray [N,4] -> keypoints ndarray [N, num_kpts, 3]."""
    if flip_test:
        raise NotImplementedError("flip_test is not on the hot path (the reference accepts the flag and ignores it, net_utils.py:36)")
    boxes = np.asarray(boxes)
    n = boxes.shape[0]
    if n == 0:
        return np.zeros((0, getattr(net, "num_classes", 17), 3), np.float32)
    centers, scales = boxes_to_center_scale(boxes, inp_res)
    crops = torch.stack([transform_image(im, centers[i], scales[i], inp_res) for i in range(n)]).float()
    dev = next(net.parameters()).device
    outs = []
    for s in range(0, n, max_batch):
        e = min(n, s + max_batch)
        hm = net(crops[s:e].to(dev, non_blocking=True))
        c = torch.from_numpy(centers[s:e].astype(np.float32)).to(dev)
        sc = torch.from_numpy(scales[s:e].astype(np.float32)).to(dev)
        outs.append(final_preds_cuda(hm, c, sc, True))
    return torch.cat(outs).cpu().numpy()


@torch.no_grad()
def flow_est(net, prev_im, cur_im):
    """prev_im, cur_im: HxWx3 ndarrays in BGR order -> flow ndarray [2,H,W] (u, v) in pixels."""
    prev_im, cur_im = np.flip(prev_im, axis=2), np.flip(cur_im, axis=2)
    imgs = np.array([[prev_im, cur_im]]).transpose((0, 4, 1, 2, 3)).astype(np.float32)      # [1, 3(RGB), 2(pair), H, W]
    dev = next(net.parameters()).device
    return net(torch.from_numpy(np.ascontiguousarray(imgs)).to(dev)).cpu().squeeze(0).numpy()
