"""Time corr_tc_kernel at [32,256,48,64] (dev tool; honours FT_DEV_SKIP).  Usage: corr_skip.py [split|fast]"""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from flowtrack.pytorch_b200 import _lib, engine  # noqa: E402

prec = _lib.PREC_FAST if (len(sys.argv) > 1 and sys.argv[1] == "fast") else _lib.PREC_SPLIT
B, C, H, W = int(os.environ.get("CORR_B", "32")), 256, 48, 64
L = _lib.lib()
g = torch.Generator(device="cuda").manual_seed(0)
acts = []
for i in range(2):
    t = torch.randn(B, C, H, W, device="cuda", generator=g)
    act = engine.Act(B, H, W, C, "cuda")
    st = t.stride()
    _lib.check(L.ft_pack_image(t.data_ptr(), st[0], st[1], st[2], st[3], C, 1.0, act.struct(), 0, 0, 0), "pack")
    acts.append(act)
if os.environ.get("OUT_ALLOC") == "big":      # the output as a 2 MiB-aligned slice of one 2 GiB allocation (page-size / TLB experiments)
    pool = torch.empty(2 << 30, dtype=torch.uint8, device="cuda")
    off = (-pool.data_ptr()) % (2 << 20)
    o = pool[off:off + B * 441 * H * W * 4].view(torch.float32).view(B, 441, H, W)
else:
    o = torch.empty(B, 441, H, W, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
flush2 = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")
clean = os.environ.get("FLUSH_CLEAN", "0") == "1"
ts = []
for i in range(8):
    flush.zero_()
    if clean:
        flush2.sum()      # read 256 MB: L2 ends up holding clean lines instead of the memset's dirty ones
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if os.environ.get("CORR_PLANAR") == "1":
        _lib.check(L.ft_correlation_act_planar(acts[0].struct(), acts[1].struct(), o.data_ptr(), 1, prec, 0), "corr")
    else:
        _lib.check(L.ft_correlation_act(acts[0].struct(), acts[1].struct(), o.data_ptr(), 1, prec, 0), "corr")
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print("B=%d parts=%s clean=%d skip=%s %s: %.1f us (median of %d, min %.1f)" % (B, os.environ.get("FT_DEV_CORR_PARTS", "auto"), clean, os.environ.get("FT_DEV_SKIP", "0"), sys.argv[1] if len(sys.argv) > 1 else "split",
                                                       statistics.median(ts[2:]) * 1e3, len(ts) - 2, min(ts) * 1e3))
if int(os.environ.get("FT_DEV_SKIP", "0")) & 64:
    import ctypes
    import numpy as np
    buf = np.zeros((148, 16), np.int64)
    L.ft_dev_corr_profile.argtypes = [ctypes.c_void_p]
    L.ft_dev_corr_profile(buf.ctypes.data)
    m = buf.astype(np.float64)
    names = {0: "MMA      [wait A-full, wait acc-empty, wait B-full, total]", 1: "producer [wait empty, -, -, total]", 2: "epilogue0 [wait acc-full, -, -, total]"}
    for r in range(3):
        if r == 0:
            print("  MMA extras [fence, issue block (elect..commit..syncwarp), commits] mean kclk:", np.round(m[:, 12:15].mean(0) / 1e3, 1))
        print("  %-62s mean kclk: %s   (max total %.0f)" % (names[r], np.round(m[:, r * 4:r * 4 + 4].mean(0) / 1e3, 1), m[:, r * 4 + 3].max() / 1e3))
