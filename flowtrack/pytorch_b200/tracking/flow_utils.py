"""box_propagation with the reference signature (lib/tracking/flow_utils.py:7-35), computed by
ft_box_propagation on the GPU.

Restated semantics (same as oracle/decode_oracle.py): H, W are the flow field's spatial dims (the
reference reads shape[0]/shape[1] of a [2,H,W] array, :19-20); the flow is looked up at the truncated
key-point pixel (the reference's list-of-lists index means flow[xs, ys] on the (2,1,0) transpose,
:21-24).  One deliberate difference: key-point pixels are clamped into the flow field before the look-up
(the reference indexes unclamped and raises / wraps on out-of-image key-points)."""
import numpy as np
import torch

from .. import _lib, engine


def box_propagation_cuda(keypoints, flow, frame_of=None, out=None):
    """keypoints: CUDA fp32 [N,K,3]; flow: CUDA fp32 [B,2,H,W] (or [2,H,W]); frame_of: CUDA int32 [N] (None = all
    people belong to pair 0).  Returns CUDA fp32 boxes [N,4].  Enqueues one kernel; no sync."""
    if flow.dim() == 3:
        flow = flow.unsqueeze(0)
    if not (keypoints.is_cuda and flow.is_cuda):
        raise RuntimeError("box_propagation_cuda needs CUDA tensors (no CPU path)")
    assert keypoints.dtype == torch.float32 and flow.dtype == torch.float32 and keypoints.is_contiguous() and flow.is_contiguous()
    N, K, _ = keypoints.shape
    B, _, H, W = flow.shape
    with torch.cuda.device(keypoints.device):
        if out is None:
            out = torch.empty(N, 4, device=keypoints.device, dtype=torch.float32)
        if N == 0:
            return out
        _lib.check(_lib.lib().ft_box_propagation(keypoints.data_ptr(), frame_of.data_ptr() if frame_of is not None else None, N, K,
                                                 flow.data_ptr(), B, H, W, out.data_ptr(), engine.stream_ptr()), "box_propagation")
    return out


def box_propagation(keypoints, flow):
    """Reference signature: keypoints ndarray [num_people, num_keypoints, 3] (x, y, score), flow ndarray [2,H,W]
    -> boxes ndarray [num_people, 4]."""
    kp = torch.as_tensor(np.ascontiguousarray(keypoints, dtype=np.float32)).cuda()
    fl = torch.as_tensor(np.ascontiguousarray(flow, dtype=np.float32)).cuda()
    return box_propagation_cuda(kp, fl).cpu().numpy()
