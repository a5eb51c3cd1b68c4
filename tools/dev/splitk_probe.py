"""What a split-K work item costs (dev tool): single layers under (block_n, pair, splitk) with parts of the split-K epilogue switched
off through ft_set_debug(FT_DBG_SKIP): 16 = no partial-sum stores, 32 = every thread fences (the old protocol), 128 = the last
arriver reads part 0 only.  Timed as a CUDA graph of 8 back-to-back launches (PDL on), like Plan.autotune()."""
import ctypes
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from flowtrack.pytorch_b200 import _lib, engine  # noqa: E402
from tests import gpu_util  # noqa: E402

LAYERS = [
    ("l3 3x3 256->256 @16x12 B64", lambda: nn.Conv2d(256, 256, 3, 1, 1, bias=False), (64, 256, 16, 12)),
    ("l3 1x1 1024->256 @16x12 B64", lambda: nn.Conv2d(1024, 256, 1, bias=False), (64, 1024, 16, 12)),
    ("l4 3x3 512->512 @8x6 B64", lambda: nn.Conv2d(512, 512, 3, 1, 1, bias=False), (64, 512, 8, 6)),
    ("l3 3x3 256->256 @16x12 B8", lambda: nn.Conv2d(256, 256, 3, 1, 1, bias=False), (8, 256, 16, 12)),
]
L = _lib.lib()
torch.manual_seed(0)
for name, ctor, shape in LAYERS:
    layer = engine.conv_from_modules(ctor(), None, _lib.ACT_RELU, device="cuda")
    xa = gpu_util.pack_nchw(torch.randn(*shape, device="cuda"), layer.cin_p, xpad=layer.in_xpad)
    ho, wo = layer.out_hw(shape[2], shape[3])
    out = engine.Act(shape[0], ho, wo, layer.cout, "cuda")
    for bn, pair in ((128, 1), (64, 1)):
        items, cta_tiles = layer.tile_counts(shape[0], shape[2], shape[3], bn, pair)
        for sk in (0, 2, 3, 4, 6):
            ws = (torch.empty(cta_tiles * max(sk, 1) * bn * 128 * 4, dtype=torch.uint8, device="cuda"), torch.zeros(cta_tiles, dtype=torch.int32, device="cuda"))
            row = []
            for dbg in ((0,) if sk == 0 else (0, 32, 16, 128, 16 | 128)):
                L.ft_set_debug(0, dbg)
                d = layer.desc(xa, out, None, None, tiling=(bn, pair, 1, sk), splitk_ws=ws)
                fn = lambda: _lib.check(L.ft_conv2d_fprop(ctypes.byref(d), gpu_util.sp()), "conv")
                fn()
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for _ in range(8):
                        fn()
                g.replay()
                torch.cuda.synchronize()
                ts = []
                for _ in range(5):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    g.replay()
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1) * 1000 / 8)
                row.append("dbg%-3d %6.1f" % (dbg, min(ts)))
                ws[1].zero_()
                L.ft_set_debug(0, 0)
            print("%-30s BN=%-3d pair=%d items=%-4d splitk=%d : %s" % (name, bn, pair, items * max(sk, 1), sk, " | ".join(row)), flush=True)
