"""oracle/flowvis_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatement of the reference's Middlebury flow colour coding, lib/flownet/utils/flowlib.py:
  make_color_wheel  :474-521   55-entry RY/YG/GC/CB/BM/MR wheel
  compute_color     :430-471   angle -> wheel interpolation, saturation from the radius
  flow_to_image     :243-282   unknown-flow mask (|u|,|v| > 1e7 -> 0), normalisation by the largest radius
(the step after the flow forward in tools/flownet/demo.py:92-99).  Pinned by tests/test_oracle_golden.py against fixtures
produced by the reference's own flowlib.flow_to_image (tests/golden/make_golden.py, keys flowvis_*).  Only tests/ may
import this module."""
import numpy as np

UNKNOWN_FLOW_THRESH = 1e7


def make_color_wheel():
    RY, YG, GC, CB, BM, MR = 15, 6, 4, 11, 13, 6
    wheel = np.zeros([RY + YG + GC + CB + BM + MR, 3])
    col = 0
    wheel[0:RY, 0] = 255
    wheel[0:RY, 1] = np.floor(255 * np.arange(0, RY) / RY)
    col += RY
    wheel[col:col + YG, 0] = 255 - np.floor(255 * np.arange(0, YG) / YG)
    wheel[col:col + YG, 1] = 255
    col += YG
    wheel[col:col + GC, 1] = 255
    wheel[col:col + GC, 2] = np.floor(255 * np.arange(0, GC) / GC)
    col += GC
    wheel[col:col + CB, 1] = 255 - np.floor(255 * np.arange(0, CB) / CB)
    wheel[col:col + CB, 2] = 255
    col += CB
    wheel[col:col + BM, 2] = 255
    wheel[col:col + BM, 0] = np.floor(255 * np.arange(0, BM) / BM)
    col += BM
    wheel[col:col + MR, 2] = 255 - np.floor(255 * np.arange(0, MR) / MR)
    wheel[col:col + MR, 0] = 255
    return wheel


def compute_color(u, v):
    h, w = u.shape
    img = np.zeros([h, w, 3])
    nan = np.isnan(u) | np.isnan(v)
    u = np.where(nan, 0, u)
    v = np.where(nan, 0, v)
    wheel = make_color_wheel()
    ncols = wheel.shape[0]
    rad = np.sqrt(u ** 2 + v ** 2)
    a = np.arctan2(-v, -u) / np.pi
    fk = (a + 1) / 2 * (ncols - 1) + 1
    k0 = np.floor(fk).astype(int)
    k1 = k0 + 1
    k1[k1 == ncols + 1] = 1
    f = fk - k0
    for i in range(3):
        tmp = wheel[:, i]
        col0, col1 = tmp[k0 - 1] / 255, tmp[k1 - 1] / 255
        col = (1 - f) * col0 + f * col1
        idx = rad <= 1
        col[idx] = 1 - rad[idx] * (1 - col[idx])
        col[~idx] *= 0.75
        img[:, :, i] = np.uint8(np.floor(255 * col * (1 - nan)))
    return img


def flow_to_image(flow):
    """flow: [H,W,2] (u, v) -> uint8 [H,W,3] (RGB).  Arithmetic follows numpy's promotion on the reference's expressions for a
    float32 flow: radius in float32, normalised components in float64 (maxrad + np.finfo(float).eps is a float64 scalar)."""
    flow = np.array(flow, copy=True)
    u, v = flow[:, :, 0], flow[:, :, 1]
    unknown = (abs(u) > UNKNOWN_FLOW_THRESH) | (abs(v) > UNKNOWN_FLOW_THRESH)
    u[unknown] = 0
    v[unknown] = 0
    rad = np.sqrt(u ** 2 + v ** 2)
    maxrad = max(-1, np.max(rad))
    u = u / (maxrad + np.finfo(float).eps)
    v = v / (maxrad + np.finfo(float).eps)
    img = compute_color(u, v)
    img[np.repeat(unknown[:, :, np.newaxis], 3, axis=2)] = 0
    return np.uint8(img)
