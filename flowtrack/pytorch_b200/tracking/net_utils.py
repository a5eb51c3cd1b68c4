"""detect / pose_est / flow_est with the reference signatures (lib/tracking/net_utils.py:14-92).  The Faster R-CNN
detector itself is out of scope (person boxes are an input of the hot path): `detect` takes any callable with the
contract of the reference's `im_detect` and does the part that is on the path -- merging the propagated boxes and the NMS.

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


def final_preds_cuda(heatmaps, centers, scales, adjust_coords=True, out=None, threshold=0.0, valid=None):
    """max_preds + final_preds + transform_preds (lib/pose/utils/evaluation.py:11-35) in one kernel.
    heatmaps: CUDA fp32 [N,C,h,w]; centers [N,2], scales [N] CUDA fp32 -> CUDA fp32 [N,C,3] = (x, y, score).
    `threshold` is max_preds' argument (:11,18): heat-map coordinates are zeroed unless score > threshold."""
    assert heatmaps.is_cuda and heatmaps.dtype == torch.float32 and heatmaps.is_contiguous()
    N, C, h, w = heatmaps.shape
    with torch.cuda.device(heatmaps.device):
        if out is None:
            out = torch.empty(N, C, 3, device=heatmaps.device, dtype=torch.float32)
        _lib.check(_lib.lib().ft_decode_keypoints(heatmaps.data_ptr(), N, C, h, w, centers.data_ptr(), scales.data_ptr(), int(bool(adjust_coords)),
                                                  float(threshold), valid.data_ptr() if valid is not None else None, out.data_ptr(),
                                                  engine.stream_ptr()), "decode_keypoints")
    return out


def crop_boxes_cuda(frames, centers, scales, frame_of=None, inp_res=(256, 192), reverse_channels=False, out=None):
    """All person crops of a batch in one launch (ft_crop_affine; transform_image / cv2.warpAffine semantics).
    frames: CUDA uint8 [F,H,W,C] (cv2 layout) or float32 [F,C,H,W] / [F,H,W,C]-strided view; centers [N,2], scales [N]
    CUDA fp32; frame_of [N] int32 or None (all crops from frame 0) -> CUDA fp32 [N,C,res_h,res_w]."""
    if not frames.is_cuda:
        raise RuntimeError("crop_boxes_cuda needs CUDA tensors (no CPU path)")
    is_u8 = frames.dtype == torch.uint8
    assert is_u8 or frames.dtype == torch.float32
    if is_u8:                                    # [F,H,W,C]
        F_, H, W, C = frames.shape
        sn, sh, sw, sc = frames.stride()
    else:                                        # [F,C,H,W]
        F_, C, H, W = frames.shape
        sn, sc, sh, sw = frames.stride()
    N = centers.shape[0]
    with torch.cuda.device(frames.device):
        if out is None:
            out = torch.empty(N, C, inp_res[0], inp_res[1], device=frames.device, dtype=torch.float32)
        _lib.check(_lib.lib().ft_crop_affine(frames.data_ptr(), int(is_u8), sn, sc, sh, sw, H, W, C, centers.data_ptr(), scales.data_ptr(),
                                             frame_of.data_ptr() if frame_of is not None else None, N, inp_res[0], inp_res[1],
                                             int(bool(reverse_channels)), out.data_ptr(), engine.stream_ptr()), "crop_affine")
    return out


def nms_cuda(dets, thresh, counts=None, inclusive=True):
    """Greedy NMS of every frame in one launch (ft_nms).  dets: CUDA fp32 [F,N,5]; counts: CUDA int32 [F] or None
    -> (keep int32 [F,N] original indices in descending score order, -1 padded; num_keep int32 [F])."""
    if not dets.is_cuda:
        raise RuntimeError("nms_cuda needs CUDA tensors (no CPU path)")
    assert dets.dtype == torch.float32 and dets.is_contiguous() and dets.dim() == 3 and dets.shape[2] == 5
    F_, N, _ = dets.shape
    with torch.cuda.device(dets.device):
        keep = torch.empty(F_, N, dtype=torch.int32, device=dets.device)
        num = torch.empty(F_, dtype=torch.int32, device=dets.device)
        _lib.check(_lib.lib().ft_nms(dets.data_ptr(), counts.data_ptr() if counts is not None else None, F_, N, float(thresh), int(bool(inclusive)),
                                     keep.data_ptr(), num.data_ptr(), engine.stream_ptr()), "nms")
    return keep, num


def nms(dets, thresh):
    """detection.model.nms_wrapper.nms / pth_nms signature (lib/detection/nms/pth_nms.py:5-46): dets tensor [N,5] ->
    LongTensor of kept indices, highest score first.  A CPU tensor takes the reference's cpu_nms rule (IoU >= thresh,
    src/nms.c:57), a CUDA tensor its gpu rule (IoU > thresh, nms_kernel.cu:62); both run on the GPU here."""
    inclusive = not dets.is_cuda
    d = dets.detach().float().contiguous().cuda().unsqueeze(0)
    if d.shape[1] == 0:
        return torch.zeros(0, dtype=torch.long, device=dets.device)
    keep, num = nms_cuda(d, thresh, inclusive=inclusive)
    return keep[0, :int(num[0].item())].long().to(dets.device)


def detect(net, im, person_id, thresh=0.3, prop_dets=None):
    """lib/tracking/net_utils.py:14-34.  net: callable im -> (scores [R, num_classes], boxes [R, 4*num_classes]) (the contract
    of detection.model.test.im_detect; the detector is not part of this package) or None (no detector: only the propagated
    boxes survive).  Returns dets ndarray [N,5] after NMS over detected + propagated person boxes (GPU, ft_nms)."""
    if net is not None:
        scores, boxes = net(im)
        scores, boxes = np.asarray(scores), np.asarray(boxes)
        dets = np.hstack((boxes[:, 4 * person_id:4 * (person_id + 1)], scores[:, person_id][:, np.newaxis])).astype(np.float32)
    else:
        dets = np.zeros((0, 5), np.float32)
    if prop_dets is not None:                   # the reference writes `prop_dets != None` (elementwise on arrays)
        dets = np.concatenate((dets, np.asarray(prop_dets, np.float32)), axis=0)
    if dets.shape[0] == 0:
        return dets
    keep = nms(torch.from_numpy(dets), thresh)
    return dets[keep.numpy(), :]


@torch.no_grad()
def pose_est(net, im, boxes, inp_res=(256, 192), max_batch=32, flip_test=False):
    """im: HxWx3 ndarray (BGR, as loaded by cv2); boxes: ndarray [N,4] -> keypoints ndarray [N, num_kpts, 3].
    The frame is uploaded once; crops, network and decode run on the GPU.  The reference accepts `flip_test` and ignores
    it (net_utils.py:36); here flip_test=True runs the flip test of tools/pose/main.py:289-299 (mirrored crops in the same
    batch, left/right joints swapped back, heat-maps averaged) through evaluation.flip_test_forward."""
    boxes = np.asarray(boxes)
    n = boxes.shape[0]
    if n == 0:
        return np.zeros((0, getattr(net, "num_classes", 17), 3), np.float32)
    centers, scales = boxes_to_center_scale(boxes, inp_res)
    dev = next(net.parameters()).device
    frame = torch.from_numpy(np.ascontiguousarray(im)).to(dev)
    if frame.dtype != torch.uint8:
        frame = frame.float().permute(2, 0, 1).contiguous()          # [C,H,W] fp32
    frame = frame.unsqueeze(0)
    c_all = torch.from_numpy(centers.astype(np.float32)).to(dev)
    s_all = torch.from_numpy(scales.astype(np.float32)).to(dev)
    outs = []
    for s in range(0, n, max_batch):
        e = min(n, s + max_batch)
        crops = crop_boxes_cuda(frame, c_all[s:e].contiguous(), s_all[s:e].contiguous(), None, inp_res)
        if flip_test:
            from ..pose.utils.evaluation import flip_test_forward
            hm = flip_test_forward(net, crops)
        else:
            hm = net(crops)
        outs.append(final_preds_cuda(hm, c_all[s:e].contiguous(), s_all[s:e].contiguous(), True))
    return torch.cat(outs).cpu().numpy()


@torch.no_grad()
def flow_est(net, prev_im, cur_im):
    """prev_im, cur_im: HxWx3 ndarrays in BGR order -> flow ndarray [2,H,W] (u, v) in pixels."""
    prev_im, cur_im = np.flip(prev_im, axis=2), np.flip(cur_im, axis=2)
    imgs = np.array([[prev_im, cur_im]]).transpose((0, 4, 1, 2, 3)).astype(np.float32)      # [1, 3(RGB), 2(pair), H, W]
    dev = next(net.parameters()).device
    return net(torch.from_numpy(np.ascontiguousarray(imgs)).to(dev)).cpu().squeeze(0).numpy()
