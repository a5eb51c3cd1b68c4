"""FlowNetC (lib/flownet/networks/FlowNetC.py:13-128): siamese conv1-3, Correlation(pad 20, k 1, d 20, s1 1, s2 2)
+ LeakyReLU, 32-channel redirect, 473-channel merge, then the FlowNetS tail."""
import torch.nn as nn

from ... import _lib
from . import submodules as sm
from .FlowNetS import encoder_tail
from .correlation_package.modules.correlation import Correlation


class FlowNetC(nn.Module):
    def __init__(self, args=None, batchNorm=True, div_flow=20):
        super().__init__()
        bn = self.batchNorm = batchNorm
        self.div_flow = div_flow
        self.conv1, self.conv2, self.conv3 = sm.conv(bn, 3, 64, 7, 2), sm.conv(bn, 64, 128, 5, 2), sm.conv(bn, 128, 256, 5, 2)
        self.conv_redir = sm.conv(bn, 256, 32, kernel_size=1, stride=1)
        self.corr = Correlation(pad_size=20, kernel_size=1, max_displacement=20, stride1=1, stride2=2, corr_multiply=1)
        self.corr_activation = nn.LeakyReLU(0.1, inplace=True)
        self.conv3_1, self.conv4, self.conv4_1 = sm.conv(bn, 473, 256), sm.conv(bn, 256, 512, stride=2), sm.conv(bn, 512, 512)
        self.conv5, self.conv5_1 = sm.conv(bn, 512, 512, stride=2), sm.conv(bn, 512, 512)
        self.conv6, self.conv6_1 = sm.conv(bn, 512, 1024, stride=2), sm.conv(bn, 1024, 1024)
        sm.add_decoder_modules(self, bn, up_bias=True, inter=False)
        sm.flow_init(self)
        self.upsample1 = nn.Upsample(scale_factor=4, mode="bilinear")


def build(p, L, frames, B, tag=""):
    """frames: split-fp16 NHWC [2B,H,W,8]: images 0..B-1 = first frames, B..2B-1 = second frames (3 -> 8 channels)."""
    H, W = frames.h, frames.w
    lib = _lib.lib()
    cat = sm.alloc_concats(p, B, H, W)
    t1 = p.conv(L["conv1"], frames, p.act(2 * B, H // 2, W // 2, 64))               # both streams in one launch
    c2a = p.conv(L["conv2"], t1.view(0, 64, 0, B), cat[2].view(0, 128))             # top stream feeds concat2 (FlowNetC.py:123)
    c2b = p.conv(L["conv2"], t1.view(0, 64, B, B), p.act(B, H // 4, W // 4, 128))
    h8, w8 = H // 8, W // 8
    merged = p.act(B, h8, w8, 32 + 441, zero=True)                                  # [conv_redir(32) | corr(441 -> 448)]
    if lib.ft_correlation_tc_supported(256, h8, w8, 20, 1, 20, 1, 2) and h8 % 2 == 0 and w8 % 2 == 0:
        # tcgen05 banded GEMM straight on the split-fp16 activations of the two conv3 streams (ft_corr.cu), LeakyReLU fused;
        # parity-planar fp32 in between (dense 128-byte warp stores), re-laid out into the merge buffer by ft_pack_corr_planar
        c3a = p.conv(L["conv3"], c2a, p.act(B, h8, w8, 256))
        c3b = p.conv(L["conv3"], c2b, p.act(B, h8, w8, 256))
        corr = p.f32(B, 441, 4, h8 // 2, w8 // 2)
        p.call("correlation", lib.ft_correlation_act_planar, c3a.struct(), c3b.struct(), corr.data_ptr(), 1, p.precision)
        p.call("pack_corr", lib.ft_pack_corr_planar, corr.data_ptr(), merged.view(32, 441).struct())
    else:                                                                          # frames wider than 512 px: CUDA-core kernel on fp32 copies
        corr = p.f32(B, 441, h8, w8)
        a3f, b3f = p.f32(B, 256, h8, w8), p.f32(B, 256, h8, w8)
        c3a = p.conv(L["conv3"], c2a, p.act(B, h8, w8, 256), out_f32=a3f)
        p.conv(L["conv3"], c2b, None, out_f32=b3f)
        p.call("correlation", lib.ft_correlation_forward, a3f.data_ptr(), b3f.data_ptr(), corr.data_ptr(), B, 256, h8, w8, 20, 1, 20, 1, 2, 1)
        st = corr.stride()
        p.call("pack_image", lib.ft_pack_image, corr.data_ptr(), st[0], st[1], st[2], st[3], 441, 1.0, merged.view(32, 441).struct(), 0, 0)
    p.conv(L["conv_redir"], c3a, merged.view(0, 32))
    p.taps[tag + "corr"] = corr
    c3 = p.conv(L["conv3_1"], merged, cat[3].view(0, 256))
    c6 = encoder_tail(p, L, cat, c3, B)
    return sm.decoder(p, L, cat, c6, B, inter=False, tag=tag)
