"""oracle/decode_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatement of the host-side decode / propagation steps of the reference:

  max_preds        lib/pose/utils/evaluation.py:11-20
  final_preds      lib/pose/utils/evaluation.py:22-35
  get_transform / transform_point / transform_preds
                   lib/pose/utils/transforms.py:173-228   (rot = 0, factor = 1 branch)
  box_propagation  lib/tracking/flow_utils.py:7-35
  pose_est box->(center, scale)   lib/tracking/net_utils.py:49-50

Three documented deviations from the literal source, each forced by API drift since torch 0.4 /
numpy 1.x; everything else is restated as written:
  1. `y = indexes.div(w)` was integer division in torch 0.4 -> restated as `idx // w`.
  2. `H = flow.shape[0]; W = flow.shape[1]` on a [2,H,W] array reads (2, H) -> restated with the
     intended `H = flow.shape[1]; W = flow.shape[2]`.
  3. `flow[pos]` with a list-of-lists index meant `flow[xs, ys]` on the (2,1,0)-transposed array ->
     restated as `flow_t[xs, ys]`.  As in the reference, key-points are NOT clamped; callers must
     keep them inside the image (the CUDA kernel clamps and says so).
"""
import numpy as np


def max_preds(heatmap, threshold=0.0):
    hm = np.asarray(heatmap, dtype=np.float32)
    n, c, h, w = hm.shape
    flat = hm.reshape(n, c, -1)
    idx = flat.argmax(-1)                      # first maximum wins, like torch.max
    scores = np.take_along_axis(flat, idx[..., None], -1)
    coords = np.stack((idx % w, idx // w), axis=2).astype(np.float32)
    coords *= (scores > threshold).astype(np.float32)
    return coords, scores.astype(np.float32)


def get_transform(center, scale, res):
    t = np.eye(3)
    t[0, 0] = t[1, 1] = res[0] / scale
    t[0, 2] = -res[0] * center[0] / scale + 0.5 * res[1]
    t[1, 2] = -res[0] * center[1] / scale + 0.5 * res[0]
    return t


def transform_preds(coords, center, scale, res):
    out = np.array(coords, dtype=np.float32, copy=True)
    for i in range(out.shape[0]):
        tinv = np.linalg.inv(get_transform(center[i], scale[i], res))
        pts = np.concatenate((out[i], np.ones((out.shape[1], 1))), axis=1)
        out[i] = np.dot(tinv, pts.T)[:2].T
    return out


def final_preds(heatmap, center, scale, adjust_coords=False):
    hm = np.asarray(heatmap, dtype=np.float32)
    coords, scores = max_preds(hm)
    n, c, h, w = hm.shape
    if adjust_coords:
        for i in range(n):
            for j in range(c):
                x, y = int(coords[i, j, 0]), int(coords[i, j, 1])
                if 0 < x < w - 1 and 0 < y < h - 1:
                    diff = np.array([hm[i, j, y, x + 1] - hm[i, j, y, x - 1], hm[i, j, y + 1, x] - hm[i, j, y - 1, x]])
                    coords[i, j] += np.sign(diff) * 0.25
    return transform_preds(coords, center, scale, (h, w)), scores


def box_propagation(keypoints, flow, clamp=False):
    """clamp=True: key-point pixels are clamped into the flow field before the look-up, as ft_box_propagation does (the
    reference indexes unclamped and raises / wraps for key-points outside the frame)."""
    kp = np.asarray(keypoints, dtype=np.float64)
    flow = np.asarray(flow, dtype=np.float64)
    extend_factor = 0.15
    H, W = flow.shape[1], flow.shape[2]
    num_kpts = kp.shape[1]
    flow_t = flow.transpose((2, 1, 0))         # [W, H, 2]
    pos = kp[:, :, :2].reshape(-1, 2).T.astype(int)
    if clamp:
        pos[0], pos[1] = np.clip(pos[0], 0, W - 1), np.clip(pos[1], 0, H - 1)
    offset = flow_t[pos[0], pos[1]].reshape(-1, num_kpts, 2)
    shift = kp[:, :, :2] + offset
    mask = (kp[:, :, 2] > 0)[:, :, np.newaxis]
    min_ = np.min(shift + (1 - mask) * max(H, W), axis=1)
    max_ = np.max(shift * mask, axis=1)
    extend = (max_ - min_) * extend_factor / 2
    up_left = np.fmax(min_ - extend, 0)
    bottom_right = np.fmin(max_ + extend, np.array([W - 1, H - 1]))
    return np.concatenate((up_left, bottom_right), axis=1)


def boxes_to_center_scale(boxes, inp_res=(256, 192)):
    boxes = np.asarray(boxes, dtype=np.float64)
    centers = np.stack((boxes[:, [0, 2]].mean(1), boxes[:, [1, 3]].mean(1)), axis=1)
    scales = np.maximum(boxes[:, 3] - boxes[:, 1], (boxes[:, 2] - boxes[:, 0]) / inp_res[1] * inp_res[0])
    return centers, scales
