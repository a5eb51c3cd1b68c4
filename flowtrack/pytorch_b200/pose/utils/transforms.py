"""Host-side affine helpers with the reference names and signatures (lib/pose/utils/transforms.py:173-240, the
factor = 1 / rot = 0 / delta = 0 branch the tracking path uses): get_transform, do_transfrom, transform_point,
transform_preds, transform_image.  Pure numpy/cv2 host glue (A11); the hot path applies the same maps on the GPU
(ft_decode_keypoints for the inverse map of the key-points, ft_crop_affine for the crops)."""
import numpy as np

from ...tracking.net_utils import get_transform, transform_image  # noqa: F401


def do_transfrom(pt, t):
    """transforms.py:197-205 (spelling as in the reference): pt [(M,) N, 2] -> t . [pt, 1]."""
    pt = np.asarray(pt)
    shape = pt.shape
    if len(shape) >= 3:
        pt = pt.reshape((-1, 2))
    pt_ = np.concatenate((pt, np.ones((pt.shape[0], 1))), axis=1)
    new_pt = np.dot(t, pt_.T)[:2].T
    if len(shape) >= 3:
        new_pt = new_pt.reshape(shape)
    return new_pt


def transform_point(pt, center, scale, res, invert=0):
    """transforms.py:208-213: pixel <-> crop coordinates of the points pt [(M,) N, 2]."""
    t = get_transform(center, scale, res)
    if invert:
        t = np.linalg.inv(t)
    return do_transfrom(pt, t)


def transform_preds(coords, center, scale, res):
    """transforms.py:216-221: coords [B, J, 2] in heat-map pixels, center [B,2], scale [B] -> image coordinates (in place)."""
    for i in range(coords.shape[0]):
        coords[i] = transform_point(coords[i], center[i], scale[i], res, invert=1)
    return coords


def get_pairs(dataset="coco"):
    """transforms.py:60-96: (joint_right, joint_left) index lists of the matched left/right joints."""
    if dataset == "mpii":
        return [0, 1, 2, 10, 11, 12], [5, 4, 3, 15, 14, 13]
    if dataset == "aic":
        return [0, 1, 2, 6, 7, 8], [3, 4, 5, 9, 10, 11]
    if dataset == "coco":
        return [2, 4, 6, 8, 10, 12, 14, 16], [1, 3, 5, 7, 9, 11, 13, 15]
    raise ValueError("Not supported dataset: " + dataset)


def swap_permutation(num_joints, dataset="coco"):
    """Channel permutation applied by swaplr_image (transforms.py:140-170): perm[c] = the channel that lands on c."""
    right, left = get_pairs(dataset)
    perm = list(range(num_joints))
    for r, l in zip(right, left):
        perm[r], perm[l] = l, r
    return perm
