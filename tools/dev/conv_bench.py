"""Time single layers under explicit tilings (dev tool): conv_bench.py  -> table of (layer, block_n, pair) timings."""
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from flowtrack.pytorch_b200 import _lib, engine  # noqa: E402
from tests import gpu_util  # noqa: E402

prec = _lib.PREC_FAST if (len(sys.argv) > 1 and sys.argv[1] == "fast") else _lib.PREC_SPLIT
LAYERS = [
    ("deconv6 256->256 @32x24 B64", lambda: nn.ConvTranspose2d(256, 256, 4, 2, 1, bias=False), (64, 256, 32, 24)),
    ("l3 3x3 256->256 @16x12 B64", lambda: nn.Conv2d(256, 256, 3, 1, 1, bias=False), (64, 256, 16, 12)),
    ("l4 3x3 512->512 @8x6 B64", lambda: nn.Conv2d(512, 512, 3, 1, 1, bias=False), (64, 512, 8, 6)),
    ("flow conv3_1 256->256 @48x64 B16", lambda: nn.Conv2d(256, 256, 3, 1, 1), (16, 256, 48, 64)),
    ("l2 1x1 512->128 @32x24 B64", lambda: nn.Conv2d(512, 128, 1, bias=False), (64, 512, 32, 24)),
    ("l3 1x1 256->1024 @16x12 B64", lambda: nn.Conv2d(256, 1024, 1, bias=False), (64, 256, 16, 12)),
    ("l3 1x1 1024->256 @16x12 B64", lambda: nn.Conv2d(1024, 256, 1, bias=False), (64, 1024, 16, 12)),
    ("l2 1x1 128->512 @32x24 B64", lambda: nn.Conv2d(128, 512, 1, bias=False), (64, 128, 32, 24)),
    ("l1 1x1 64->256 @64x48 B64", lambda: nn.Conv2d(64, 256, 1, bias=False), (64, 64, 64, 48)),
    ("l4 1x1 512->2048 @8x6 B64", lambda: nn.Conv2d(512, 2048, 1, bias=False), (64, 512, 8, 6)),
    ("flow deconv2 392->64 @48x64 B16", lambda: nn.ConvTranspose2d(392, 64, 4, 2, 1), (16, 392, 48, 64)),
    ("pose stem 7x7/2 3->64 @256x192 B64", lambda: nn.Conv2d(3, 64, 7, 2, 3, bias=False), (64, 3, 256, 192)),
    ("l1 3x3 64->64 @64x48 B64", lambda: nn.Conv2d(64, 64, 3, 1, 1, bias=False), (64, 64, 64, 48)),
]
ONLY = os.environ.get("CB_ONLY")            # e.g. "0:128:0,0:128:1" = layer index : block_n : pair
only = set(tuple(int(v) for v in item.split(":")) for item in ONLY.split(",")) if ONLY else None
REPS = int(os.environ.get("CB_REPS", "10"))
torch.manual_seed(0)
for li, (name, ctor, shape) in enumerate(LAYERS):
    if only is not None and not any(o[0] == li for o in only):
        continue
    mod = ctor()
    layer = engine.conv_from_modules(mod, None, _lib.ACT_RELU, device="cuda")
    xa = gpu_util.pack_nchw(torch.randn(*shape, device="cuda"), layer.cin_p, xpad=layer.in_xpad)
    ho, wo = layer.out_hw(shape[2], shape[3])
    out = engine.Act(shape[0], ho, wo, layer.cout, "cuda")
    fl = layer.flops(shape[0], shape[2], shape[3])
    for bn, pair in ((64, 0), (128, 0), (256, 0), (64, 1), (128, 1), (256, 1)):
        if layer.cout_pad % bn or (only is not None and (li, bn, pair) not in only):
            continue
        d = layer.desc(xa, out, None, None, prec, _lib.IMPL_TCGEN05, tiling=(bn, pair))
        import ctypes
        fn = lambda: _lib.check(_lib.lib().ft_conv2d_fprop(ctypes.byref(d), gpu_util.sp()), "conv")
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(REPS):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1000 / REPS
        print("%-36s chunk=%-5s BN=%-3d pair=%d  %8.1f us  %6.1f TF/s alg" % (name, os.environ.get("FT_DEV_CHUNK", "4"), bn, pair, us, fl / us / 1e6), flush=True)
