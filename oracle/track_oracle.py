"""TEST INFRASTRUCTURE ONLY (oracle): CPU restatements of the tracking glue either side of the two networks --
person crops and greedy box NMS (row 2).  Only tests/, __graft_entry__.smoke() and bench.py's
CPU legs may import this.

  get_transform / crop_boxes   lib/pose/utils/transforms.py:173-184,231-240 + lib/tracking/net_utils.py:48-57
                               (transform_image -> cv2.warpAffine(img, t[:2], (res[1], res[0])), bilinear, zero border)
  warp_affine_fixed            restatement of OpenCV's bilinear warpAffine (third-party, un-vendored; this image ships
                               opencv-python 4.13.0; the reference pins no version): modules/imgproc/src/imgwarp.cpp,
                               WarpAffineInvoker + remapBilinear fixed-point path.  Pinned against cv2.warpAffine itself
                               (tests/test_oracle_golden.py: bit-exact on uint8 images) and against crops produced by the
                               reference's own transform_image (tests/golden/reference_outputs.npz, keys crop_*).
  get_pairs / swaplr_image / flip_test_merge
                               lib/pose/utils/transforms.py:60-96,140-170 and tools/pose/main.py:289-299 (flip test); pinned
                               against the reference's own swaplr_image (reference_track.npz, keys swap_*).
  compute_oks / nms_oks        lib/pose/utils/evaluation.py:60-101, restated line by line (pinned against the reference functions
                               themselves: reference_track.npz keys oks_*)
  nms_cpu                      lib/detection/nms/src/nms.c:36-62 (cpu_nms: suppress at IoU >= thresh) with the ordering
                               of pth_nms.py:13-22; inclusive=False gives the GPU variant (nms_kernel.cu:62, IoU > thresh).
                               Pinned against the reference's own cpu_nms: lib/detection/nms/src/nms.c compiles
                               unmodified against oracle/ref_shims/TH/TH.h into oracle/_ref/libref_nms.so (ref_cpu_nms
                               below drives it the way pth_nms.py:9-22 does); its outputs on seeded box sets are the
                               fixtures nms_keep_* of tests/golden/reference_track.npz.
"""
import numpy as np


def get_transform(center, scale, res):
    t = np.eye(3)
    t[0, 0] = t[1, 1] = res[0] / scale
    t[0, 2] = -res[0] * center[0] / scale + 0.5 * res[1]
    t[1, 2] = -res[0] * center[1] / scale + 0.5 * res[0]
    return t


def boxes_to_center_scale(boxes, inp_res=(256, 192)):
    boxes = np.asarray(boxes, dtype=np.float64)
    centers = np.stack((boxes[:, [0, 2]].mean(1), boxes[:, [1, 3]].mean(1)), axis=1)
    scales = np.maximum(boxes[:, 3] - boxes[:, 1], (boxes[:, 2] - boxes[:, 0]) / inp_res[1] * inp_res[0])
    return centers, scales


def warp_affine_fixed(img, M, res):
    """img: [H,W,C] uint8 or float; M: 2x3 forward map; res = (h, w) -> [h,w,C] (uint8 in -> uint8 out, else float32)."""
    M = np.asarray(M, dtype=np.float64)[:2].copy().ravel()
    D = M[0] * M[4] - M[1] * M[3]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = M[4] * D, M[0] * D
    M[0], M[1], M[3], M[4] = A11, M[1] * -D, M[3] * -D, A22
    b1 = -M[0] * M[2] - M[1] * M[5]
    b2 = -M[3] * M[2] - M[4] * M[5]
    M[2], M[5] = b1, b2
    h, w = res
    H, W = img.shape[:2]
    xs, ys = np.arange(w), np.arange(h)
    ad = np.rint(M[0] * xs * 1024).astype(np.int64)
    bd = np.rint(M[3] * xs * 1024).astype(np.int64)
    X0 = np.rint((M[1] * ys + M[2]) * 1024).astype(np.int64) + 16
    Y0 = np.rint((M[4] * ys + M[5]) * 1024).astype(np.int64) + 16
    X = (X0[:, None] + ad[None, :]) >> 5
    Y = (Y0[:, None] + bd[None, :]) >> 5
    sx, sy = np.clip(X >> 5, -32768, 32767), np.clip(Y >> 5, -32768, 32767)
    fx, fy = X & 31, Y & 31
    is_u8 = img.dtype == np.uint8
    src = img.astype(np.int64) if is_u8 else img.astype(np.float32)
    if src.ndim == 2:
        src = src[:, :, None]

    def px(yy, xx):
        ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        return src[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)] * ok[..., None]

    w00, w01, w10, w11 = (32 - fx) * (32 - fy), fx * (32 - fy), (32 - fx) * fy, fx * fy
    if is_u8:
        s = px(sy, sx) * w00[..., None] + px(sy, sx + 1) * w01[..., None] + px(sy + 1, sx) * w10[..., None] + px(sy + 1, sx + 1) * w11[..., None]
        return ((s + 512) >> 10).astype(np.uint8)
    k = np.float32(1.0 / 1024.0)
    f = lambda wt: (wt.astype(np.float32) * k)[..., None]
    return (px(sy, sx) * f(w00) + px(sy, sx + 1) * f(w01) + px(sy + 1, sx) * f(w10) + px(sy + 1, sx + 1) * f(w11)).astype(np.float32)


def crop_boxes(img, boxes, inp_res=(256, 192), use_cv2=False):
    """pose_est's crop loop: img [H,W,C] -> crops [N,C,res_h,res_w] float32 (values as warpAffine returns them)."""
    centers, scales = boxes_to_center_scale(boxes, inp_res)
    out = []
    for c, s in zip(centers, scales):
        M = get_transform(c, s, inp_res)[:2]
        if use_cv2:
            import cv2
            o = cv2.warpAffine(img, M, (inp_res[1], inp_res[0]))
            if o.ndim == 2:
                o = o[:, :, None]
        else:
            o = warp_affine_fixed(img, M, inp_res)
        out.append(o.transpose(2, 0, 1).astype(np.float32))
    return np.stack(out) if out else np.zeros((0, img.shape[2], inp_res[0], inp_res[1]), np.float32)


def nms_cpu(dets, thresh, inclusive=True):
    """dets: [N,5] float32 (x1,y1,x2,y2,score) -> indices of the kept boxes, highest score first."""
    dets = np.asarray(dets, dtype=np.float32)
    n = dets.shape[0]
    if n == 0:
        return np.zeros((0,), np.int64)
    x1, y1, x2, y2, sc = (dets[:, i] for i in range(5))
    areas = (x2 - x1 + np.float32(1)) * (y2 - y1 + np.float32(1))
    order = np.lexsort((np.arange(n), -sc.astype(np.float64)))          # descending score, ties by index
    suppressed = np.zeros(n, bool)
    keep = []
    one, zero = np.float32(1), np.float32(0)
    for _i in range(n):
        i = order[_i]
        if suppressed[i]:
            continue
        keep.append(i)
        for _j in range(_i + 1, n):
            j = order[_j]
            if suppressed[j]:
                continue
            w = max(zero, min(x2[i], x2[j]) - max(x1[i], x1[j]) + one)
            h = max(zero, min(y2[i], y2[j]) - max(y1[i], y1[j]) + one)
            inter = np.float32(w * h)
            ovr = inter / (areas[i] + areas[j] - inter)
            if (ovr >= thresh) if inclusive else (ovr > thresh):
                suppressed[j] = True
    return np.asarray(keep, np.int64)


def get_pairs(dataset="coco"):
    if dataset == "mpii":
        return [0, 1, 2, 10, 11, 12], [5, 4, 3, 15, 14, 13]
    if dataset == "aic":
        return [0, 1, 2, 6, 7, 8], [3, 4, 5, 9, 10, 11]
    if dataset == "coco":
        return [2, 4, 6, 8, 10, 12, 14, 16], [1, 3, 5, 7, 9, 11, 13, 15]
    raise ValueError(dataset)


def swaplr_image(img, dataset="coco"):
    """img [B,K,h,w]: mirror the width axis and exchange the left/right joint channels (transforms.py:140-170)."""
    right, left = get_pairs(dataset)
    img = np.asarray(img)[:, :, :, ::-1].copy()
    tmp = img[:, right].copy()
    img[:, right] = img[:, left]
    img[:, left] = tmp
    return img


def flip_test_merge(output, output_flipped, dataset="coco"):
    """tools/pose/main.py:296-299: (output + swaplr_image(output_flipped)) * 0.5"""
    return (np.asarray(output) + swaplr_image(output_flipped, dataset)) * np.float32(0.5)


def compute_oks(pred, anno, ref_scale, delta, ground_truth=True, threshold=0):
    """evaluation.py:60-81."""
    pred, anno = np.asarray(pred), np.asarray(anno)
    batch_size = pred.shape[0]
    if anno.ndim < 3:
        anno = np.tile(anno[np.newaxis, :, :], (batch_size, 1, 1))
    oks = np.zeros((batch_size))
    for i in range(batch_size):
        pj, aj, scale = pred[i], anno[i], ref_scale[i]
        if ground_truth:
            is_count = aj[:, 2] > 0
        else:
            is_count = np.logical_and(aj[:, 2] >= threshold, pj[:, 2] >= threshold)
        if is_count.sum() != 0:
            dist = pow((aj[is_count, :2] - pj[is_count, :2]), 2).sum(1)
            d = pow(delta[is_count], 2)
            oks[i] = np.exp(-dist / 2 / d / (scale + np.spacing(1))).mean()
    return oks


def nms_oks(pred, oks_thresh, delta, kpt_thresh=0):
    """evaluation.py:83-101."""
    num_preds = len(pred)
    scores = np.array([pred[i]["score"] for i in range(num_preds)])
    joints = np.array([pred[i]["joints"] for i in range(num_preds)])
    areas = np.array([pred[i]["area"] for i in range(num_preds)])
    order = scores.argsort()[::-1]
    keep = []
    while order.size > 0:
        i = order[0]
        keep.append(int(i))
        ref_scale = (areas[i] + areas[order[1:]]) / 2
        oks = compute_oks(joints[order[1:]], joints[i], ref_scale, delta, ground_truth=False, threshold=kpt_thresh)
        ind = np.where(oks <= oks_thresh)[0]
        order = order[ind + 1]
    return keep


def ref_cpu_nms(dets, thresh):
    """The reference's own cpu_nms (oracle/_ref/libref_nms.so, built by `make -C oracle ref` from the unmodified nms.c),
    called like pth_nms.py:9-22 calls it: areas = (x2-x1+1)(y2-y1+1), order = argsort(scores, descending).  Returns the kept
    indices or None when the library is not there (GPU box without /root/reference and without the prebuilt file)."""
    import ctypes
    import os
    so = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libref_nms.so")
    if not os.path.exists(so):
        return None
    lib = ctypes.CDLL(so)

    class T(ctypes.Structure):
        _fields_ = [("data", ctypes.c_void_p), ("size", ctypes.c_long * 2)]

    dets = np.ascontiguousarray(dets, dtype=np.float32)
    n = dets.shape[0]
    areas = ((dets[:, 2] - dets[:, 0] + 1) * (dets[:, 3] - dets[:, 1] + 1)).astype(np.float32)
    order = np.lexsort((np.arange(n), -dets[:, 4].astype(np.float64))).astype(np.int64)     # descending score, ties by index
    keep = np.zeros(max(n, 1), np.int64)
    num = np.zeros(1, np.int64)

    def t(a, s0, s1=1):
        return T(a.ctypes.data, (ctypes.c_long * 2)(s0, s1))

    lib.cpu_nms.argtypes = [ctypes.POINTER(T)] * 5 + [ctypes.c_float]
    lib.cpu_nms.restype = ctypes.c_int
    rc = lib.cpu_nms(ctypes.byref(t(keep, n)), ctypes.byref(t(num, 1)), ctypes.byref(t(dets, n, 5)), ctypes.byref(t(order, n)),
                     ctypes.byref(t(areas, n)), ctypes.c_float(thresh))
    assert rc == 1
    return keep[:int(num[0])].copy()


def process_frame(pose_net, flow_net, prev_im, cur_im, prev_dets, prev_keypoints, det=None, inp_res=(256, 192), thresh=0.3, max_people=None):
    """The reference's per-frame tracking step, tools/tracking/demo.py:35-42, with flow_est / detect / pose_est restated from
    lib/tracking/net_utils.py (:73-92, :25-33 minus the detector network, :36-71 with the `min` fix of :61) over CPU oracle
    networks.  prev_im, cur_im: HxWx3 uint8 BGR; prev_dets [n,5]; prev_keypoints [n,K,3]; det: optional detector boxes [d,5].
    Returns (cur_dets [m,5], cur_keypoints [m,K,3], flow [2,H,W]); max_people truncates to the highest-scored boxes (the
    capacity of the device pipeline's person slots)."""
    import torch
    from oracle import decode_oracle
    pair = np.stack([prev_im[:, :, ::-1], cur_im[:, :, ::-1]])[None].transpose(0, 4, 1, 2, 3).astype(np.float32)      # [1,3(RGB),2,H,W]
    with torch.no_grad():
        flow = flow_net(torch.from_numpy(np.ascontiguousarray(pair)))[0].numpy()
    prev_dets = np.asarray(prev_dets, np.float32).reshape(-1, 5)
    if prev_dets.shape[0]:
        prop_boxes = decode_oracle.box_propagation(np.asarray(prev_keypoints), flow, clamp=True)
        prop_dets = np.concatenate((prop_boxes, prev_dets[:, 4][:, np.newaxis]), axis=1).astype(np.float32)
    else:
        prop_dets = np.zeros((0, 5), np.float32)
    dets = np.zeros((0, 5), np.float32) if det is None else np.asarray(det, np.float32).reshape(-1, 5)
    dets = np.concatenate((dets, prop_dets), axis=0)
    if dets.shape[0]:
        dets = dets[nms_cpu(dets, thresh, inclusive=True)]
    if max_people is not None:
        dets = dets[:max_people]
    if dets.shape[0] == 0:
        return dets, np.zeros((0, 17, 3), np.float32), flow
    crops = crop_boxes(np.ascontiguousarray(cur_im), dets[:, :4], inp_res)
    with torch.no_grad():
        hm = pose_net(torch.from_numpy(crops)).numpy()
    centers, scales = boxes_to_center_scale(dets[:, :4], inp_res)
    xy, sc = decode_oracle.final_preds(hm, centers, scales, True)
    return dets, np.concatenate((xy, sc), axis=2).astype(np.float32), flow
