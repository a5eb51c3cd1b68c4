"""Helpers for the -m gpu tests: drive single layers / ops through the C ABI."""
import ctypes

import torch

from flowtrack.pytorch_b200 import _lib, engine


# The parity bar, stated once (BASELINE.json north_star): every test, smoke() and bench.py's `parity`
# key hold the CUDA path to these numbers against the oracle / the reference fixtures.
HEATMAP_TOL = 1e-3      # max-abs on heat-maps (conditioned weights: heat-map abs-max ~4)
FLOW_TOL = 1e-3         # max-abs on flows, relative to max(1, max|flow|) px -- every FlowNet2 variant
KEYPOINT_TOL_PX = 1.0   # decoded key-points / propagated boxes


def flow_tol(ref):
    return FLOW_TOL * max(1.0, float(ref.abs().max()))


def sp():
    return engine.stream_ptr()


def pack_nchw(x, cpad=None, xpad=(0, 0)):
    """fp32 NCHW cuda tensor -> split-fp16 NHWC Act via ft_pack_image (optionally with zero columns around every row)."""
    B, C, H, W = x.shape
    a = engine.Act(B, H, W, cpad or C, x.device, xpad=xpad)
    st = x.stride()
    _lib.check(_lib.lib().ft_pack_image(x.data_ptr(), st[0], st[1], st[2], st[3], C, 1.0, a.struct(), xpad[0], xpad[1], sp()), "pack_image")
    return a


def run_layer(layer, x_act, res=None, precision=_lib.PREC_SPLIT, impl=_lib.IMPL_TCGEN05, want_f32=False, out=None, tiling=None, tile_order=0):
    ho, wo = layer.out_hw(x_act.h, x_act.w)
    if out is None:
        out = engine.Act(x_act.n, ho, wo, layer.cout, x_act.hi.device, zero=True)
    f32 = torch.zeros(x_act.n, layer.cout, ho, wo, device=x_act.hi.device) if want_f32 else None
    d = layer.desc(x_act, out, f32, res, precision, impl, tiling=tiling)
    d.tile_order = tile_order
    _lib.check(_lib.lib().ft_conv2d_fprop(ctypes.byref(d), sp()), "conv2d")
    torch.cuda.synchronize()
    return (out, f32) if want_f32 else out


def strict_fp32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
