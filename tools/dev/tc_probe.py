"""First-contact probe for the tcgen05 conv kernel: a single-tile GEMM with structured error reporting,
then a few descriptor variants if the default is wrong, then early timings.  Dev tool, not a test."""
import ctypes
import os
import subprocess
import sys
import time

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from flowtrack.pytorch_b200 import _lib, engine  # noqa: E402
from tests import gpu_util  # noqa: E402


def one(cin, cout, hw, prec, impl=_lib.IMPL_TCGEN05, k=1, seed=0):
    torch.manual_seed(seed)
    conv = nn.Conv2d(cin, cout, k, 1, (k - 1) // 2, bias=False)
    x = torch.randn(1, cin, *hw)
    with torch.no_grad():
        ref = conv.double()(x.double()).float()
    layer = engine.conv_from_modules(conv.float(), None, device="cuda")
    xa = gpu_util.pack_nchw(x.cuda(), layer.cin_p)
    out = gpu_util.run_layer(layer, xa, precision=prec, impl=impl).to_nchw(cout).cpu()
    err = (out - ref).abs()
    return err, ref


def report(tag, err, ref):
    e = err.max().item()
    print("%-28s max-abs %.3e (ref max %.2f)" % (tag, e, ref.abs().max().item()), flush=True)
    if e > 1e-2:
        m = err.flatten(2)[0].t()                      # [pixels, cout]
        print("   per-row-octet err :", [round(v, 3) for v in m.view(-1, 8, m.shape[1]).amax((1, 2))[:8].tolist()])
        print("   per-col-octet err :", [round(v, 3) for v in m.view(m.shape[0], -1, 8).amax((0, 2))[:8].tolist()])
        print("   rows 0..7, col 0  :", [round(v, 3) for v in m[:8, 0].tolist()])
    return e


def main():
    print(torch.cuda.get_device_name(0), flush=True)
    e_s, ref = one(64, 128, (8, 16), _lib.PREC_SPLIT, _lib.IMPL_SIMT)
    report("simt 1x1 64->128 M=128", e_s, ref)
    bad = 0
    for tag, prec in (("tc fast", _lib.PREC_FAST), ("tc split", _lib.PREC_SPLIT)):
        for cin, cout, hw, k in ((64, 128, (8, 16), 1), (256, 128, (8, 16), 1), (64, 64, (16, 16), 3), (128, 256, (16, 24), 3)):
            try:
                err, ref = one(cin, cout, hw, prec, k=k)
                e = report("%s k%d %d->%d M=%d" % (tag, k, cin, cout, hw[0] * hw[1]), err, ref)
                bad += e > (1e-2 if prec == _lib.PREC_FAST else 1e-4)
            except Exception as ex:  # noqa
                print("%s k%d %d->%d FAILED: %s" % (tag, k, cin, cout, ex), flush=True)
                bad += 1
                return 1
    if bad and not os.environ.get("FT_DBG_DESC"):
        for var in ("64,1,2,0", "64,0,2,1", "64,1,1,1", "128,1,2,1", "8,1,2,64", "1,1,2,64"):
            r = subprocess.run([sys.executable, __file__, "--variant"], env=dict(os.environ, FT_DBG_DESC=var), capture_output=True, text=True, timeout=300)
            print("=== variant FT_DBG_DESC=%s\n%s%s" % (var, r.stdout[-1500:], r.stderr[-500:]), flush=True)
    return bad


if __name__ == "__main__":
    if "--variant" in sys.argv:
        for prec in (_lib.PREC_FAST,):
            err, ref = one(64, 128, (8, 16), prec)
            report("variant fast 64->128", err, ref)
            err, ref = one(256, 128, (8, 16), prec)
            report("variant fast 256->128", err, ref)
        sys.exit(0)
    sys.exit(1 if main() else 0)
