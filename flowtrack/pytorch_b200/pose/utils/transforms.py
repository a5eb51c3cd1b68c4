"""Host-side affine helpers with the reference names (lib/pose/utils/transforms.py:173-240): get_transform,
transform_point, transform_preds, transform_image.  Pure numpy/cv2 host glue (A11); the decode hot path applies the
same inverse affine inside ft_decode_keypoints."""
import numpy as np

from ...tracking.net_utils import get_transform, transform_image  # noqa: F401


def transform_point(pt, center, scale, res, invert=0):
    """transforms.py:204-212 (rot = 0): pixel <-> crop coordinates of one point."""
    t = get_transform(center, scale, res)
    if invert:
        t = np.linalg.inv(t)
    new_pt = np.array([pt[0], pt[1], 1.0])
    return np.dot(t, new_pt)[:2]


def transform_preds(coords, center, scale, res):
    """transforms.py:214-228: heat-map coordinates [K,2] of one person -> image coordinates."""
    coords = np.asarray(coords, dtype=np.float64).copy()
    for p in range(coords.shape[0]):
        coords[p, 0:2] = transform_point(coords[p, 0:2], center, scale, res, 1)
    return coords
