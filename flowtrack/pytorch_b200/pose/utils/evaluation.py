"""max_preds / final_preds with the reference signatures (lib/pose/utils/evaluation.py:11-35), computed by
ft_decode_keypoints on the GPU (no CPU path: heat-maps must be CUDA tensors)."""
import numpy as np
import torch

from ...tracking.net_utils import final_preds_cuda


def _as_dev(a, dev):
    return torch.as_tensor(np.ascontiguousarray(np.asarray(a, dtype=np.float32))).to(dev)


def max_preds(heatmaps, threshold=0):
    """evaluation.py:11-20: heat-maps CUDA fp32 [N,C,h,w] -> (coords ndarray [N,C,2] float32 in heat-map pixels, scores
    ndarray [N,C,1]).  y = idx // w (torch-0.4 integer division semantics); coordinates are zeroed where
    score <= threshold."""
    if not heatmaps.is_cuda:
        raise RuntimeError("max_preds: heat-maps must be CUDA tensors (there is no CPU path)")
    N, C, h, w = heatmaps.shape
    # identity transform: scale = h  ->  s = 1, centre = (w/2, h/2)
    centers = torch.tensor([[0.5 * w, 0.5 * h]], device=heatmaps.device, dtype=torch.float32).repeat(N, 1)
    scales = torch.full((N,), float(h), device=heatmaps.device, dtype=torch.float32)
    out = final_preds_cuda(heatmaps.contiguous().float(), centers, scales, adjust_coords=False, threshold=float(threshold)).cpu().numpy()
    return out[:, :, :2].copy(), out[:, :, 2:3].copy()


def final_preds(heatmaps, center, scale, adjust_coords=False):
    """evaluation.py:22-35: heat-maps CUDA fp32 [N,C,h,w], center [N,2], scale [N] (tensors or arrays)
    -> (preds ndarray [N,C,2] in image pixels, scores ndarray [N,C,1]).  Default adjust_coords=False as in the
    reference (pose_est passes True, net_utils.py:68)."""
    if not heatmaps.is_cuda:
        raise RuntimeError("final_preds: heat-maps must be CUDA tensors (there is no CPU path)")
    dev = heatmaps.device
    c = center.to(dev).float().contiguous() if torch.is_tensor(center) else _as_dev(center, dev)
    s = scale.to(dev).float().contiguous() if torch.is_tensor(scale) else _as_dev(scale, dev)
    out = final_preds_cuda(heatmaps.contiguous().float(), c.reshape(-1, 2), s.reshape(-1), adjust_coords).cpu().numpy()
    return out[:, :, :2].astype(np.float64), out[:, :, 2:3]


@torch.no_grad()
def flip_test_forward(model, inputs, dataset="coco"):
    """The flip test of tools/pose/main.py:289-299 on the GPU: output = 0.5 * (model(x) + swaplr_image(model(flip_w(x)))).
    The mirrored crops run in the SAME batch as the originals (one forward of 2B crops); ft_flip_w / ft_flip_merge do the
    mirroring, the left/right joint swap and the average.  inputs: CUDA fp32 [B,3,H,W] -> CUDA fp32 heat-maps [B,K,H/4,W/4]."""
    from ... import _lib, engine
    from .transforms import swap_permutation
    if not inputs.is_cuda:
        raise RuntimeError("flip_test_forward needs CUDA tensors (no CPU path)")
    x = inputs.contiguous().float()
    B, C, H, W = x.shape
    with torch.cuda.device(x.device):
        both = torch.empty(2 * B, C, H, W, device=x.device, dtype=torch.float32)
        both[:B].copy_(x)
        L = _lib.lib()
        _lib.check(L.ft_flip_w(x.data_ptr(), both[B:].data_ptr(), B * C * H, W, engine.stream_ptr()), "flip_w")
        hm = model(both)
        K, h, w = hm.shape[1:]
        perm = torch.tensor(swap_permutation(K, dataset), dtype=torch.int32, device=x.device)
        out = torch.empty(B, K, h, w, device=x.device, dtype=torch.float32)
        _lib.check(L.ft_flip_merge(hm[:B].data_ptr(), hm[B:].data_ptr(), perm.data_ptr(), B, K, h, w, out.data_ptr(), engine.stream_ptr()), "flip_merge")
    return out


def nms_oks(pred, oks_thresh, delta, kpt_thresh=0):
    """evaluation.py:83-101 signature: pred = list of {'score', 'joints' [K,3], 'area'} -> list of kept indices, highest
    score first.  Runs ft_nms_oks on the GPU (one launch; the reference loops in numpy).
    Order and ties: the visiting order is the reference's own `scores.argsort()[::-1]` on the float64 scores (same numpy
    call, so equal scores are visited in the reference's order); the poses are handed to the kernel in that order with
    strictly decreasing surrogate keys.  Joints and areas go to the device as fp32, the OKS itself is evaluated in
    double from them (a pose pair whose OKS sits within ~1e-7 of oks_thresh may be decided differently)."""
    from ... import _lib, engine
    n = len(pred)
    if n == 0:
        return []
    dev = torch.device("cuda")
    order = np.array([p["score"] for p in pred]).argsort()[::-1]
    joints = torch.as_tensor(np.array([pred[i]["joints"] for i in order], dtype=np.float32)).to(dev).contiguous().unsqueeze(0)
    areas = torch.as_tensor(np.array([pred[i]["area"] for i in order], dtype=np.float32)).to(dev).unsqueeze(0)
    keys = torch.arange(n, 0, -1, dtype=torch.float32, device=dev).unsqueeze(0)
    dl = torch.as_tensor(np.asarray(delta, dtype=np.float32)).to(dev)
    K = joints.shape[2]
    keep = torch.empty(1, n, dtype=torch.int32, device=dev)
    num = torch.empty(1, dtype=torch.int32, device=dev)
    _lib.check(_lib.lib().ft_nms_oks(joints.data_ptr(), keys.data_ptr(), areas.data_ptr(), dl.data_ptr(), None, 1, n, K, float(oks_thresh),
                                     float(kpt_thresh), keep.data_ptr(), num.data_ptr(), engine.stream_ptr()), "nms_oks")
    return [int(order[i]) for i in keep[0, :int(num[0].item())].cpu().tolist()]
