"""FlowNetSD (lib/flownet/networks/FlowNetSD.py:11-106): small-displacement net, 3x3 stem at full resolution and
inter-convs before every flow prediction."""
import torch.nn as nn

from . import submodules as sm
from .FlowNetS import encoder_tail


class FlowNetSD(nn.Module):
    def __init__(self, args=None, batchNorm=True):
        super().__init__()
        bn = self.batchNorm = batchNorm
        self.conv0, self.conv1, self.conv1_1 = sm.conv(bn, 6, 64), sm.conv(bn, 64, 64, stride=2), sm.conv(bn, 64, 128)
        self.conv2, self.conv2_1 = sm.conv(bn, 128, 128, stride=2), sm.conv(bn, 128, 128)
        self.conv3, self.conv3_1 = sm.conv(bn, 128, 256, stride=2), sm.conv(bn, 256, 256)
        self.conv4, self.conv4_1 = sm.conv(bn, 256, 512, stride=2), sm.conv(bn, 512, 512)
        self.conv5, self.conv5_1 = sm.conv(bn, 512, 512, stride=2), sm.conv(bn, 512, 512)
        self.conv6, self.conv6_1 = sm.conv(bn, 512, 1024, stride=2), sm.conv(bn, 1024, 1024)
        sm.add_decoder_modules(self, bn, up_bias=True, inter=True)
        sm.flow_init(self)
        self.upsample1 = nn.Upsample(scale_factor=4, mode="bilinear")


def build(p, L, x, B, tag=""):
    H, W = x.h, x.w
    cat = sm.alloc_concats(p, B, H, W)
    t = p.conv(L["conv0"], x, p.act(B, H, W, 64))
    t = p.conv(L["conv1"], t, p.act(B, H // 2, W // 2, 64))
    t = p.conv(L["conv1_1"], t, p.act(B, H // 2, W // 2, 128))
    t = p.conv(L["conv2"], t, p.act(B, H // 4, W // 4, 128))
    c2 = p.conv(L["conv2_1"], t, cat[2].view(0, 128))
    t = p.conv(L["conv3"], c2, p.act(B, H // 8, W // 8, 256))
    c3 = p.conv(L["conv3_1"], t, cat[3].view(0, 256))
    c6 = encoder_tail(p, L, cat, c3, B)
    return sm.decoder(p, L, cat, c6, B, inter=True, tag=tag)
