"""Is tcgen05 fp32 accumulation round-to-nearest or truncating?  1x1 convs with fp16-exact positive operands: every
product is exact in fp32, so the only error is the accumulation.  (dev probe, GPU box)"""
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from flowtrack.pytorch_b200 import _lib, engine  # noqa: E402
from tests import gpu_util  # noqa: E402

torch.manual_seed(0)
for K in (64, 256, 1024, 4096):
    for signed in (False, True):
        conv = nn.Conv2d(K, 128, 1, bias=False)
        with torch.no_grad():
            w = torch.rand(128, K, 1, 1) + 0.5
            x = torch.rand(2, K, 16, 16) + 0.5
            if signed:
                w = w * torch.sign(torch.randn_like(w))
                x = x * torch.sign(torch.randn_like(x))
            w, x = w.half().float(), x.half().float()
            conv.weight.copy_(w)
            ref = conv.double()(x.double())
        layer = engine.conv_from_modules(conv.float(), None, _lib.ACT_NONE, device="cuda")
        xa = gpu_util.pack_nchw(x.cuda())
        for impl, name in ((_lib.IMPL_TCGEN05, "tc  "), (_lib.IMPL_SIMT, "simt")):
            out, f32 = gpu_util.run_layer(layer, xa, precision=_lib.PREC_FAST, impl=impl, want_f32=True)
            e = (f32.cpu().double() - ref) / ref.abs().clamp_min(1e-3 * ref.abs().max())
            print("K=%4d %s %s: mean signed rel err %+.3e  max |rel err| %.3e   (K*2^-24 = %.2e)"
                  % (K, "signed  " if signed else "positive", name, e.mean().item(), e.abs().max().item(), K * 2.0 ** -24))
