"""FlowNetS (lib/flownet/networks/FlowNetS.py:15-95): parameter container + plan builder."""
import torch.nn as nn

from . import submodules as sm


class FlowNetS(nn.Module):
    def __init__(self, args=None, input_channels=12, batchNorm=True):
        super().__init__()
        bn = self.batchNorm = batchNorm
        self.input_channels = input_channels
        self.conv1, self.conv2, self.conv3 = sm.conv(bn, input_channels, 64, 7, 2), sm.conv(bn, 64, 128, 5, 2), sm.conv(bn, 128, 256, 5, 2)
        self.conv3_1, self.conv4, self.conv4_1 = sm.conv(bn, 256, 256), sm.conv(bn, 256, 512, stride=2), sm.conv(bn, 512, 512)
        self.conv5, self.conv5_1 = sm.conv(bn, 512, 512, stride=2), sm.conv(bn, 512, 512)
        self.conv6, self.conv6_1 = sm.conv(bn, 512, 1024, stride=2), sm.conv(bn, 1024, 1024)
        sm.add_decoder_modules(self, bn, up_bias=False, inter=False)
        sm.flow_init(self)
        self.upsample1 = nn.Upsample(scale_factor=4, mode="bilinear")


def encoder_tail(p, L, cat, c3, B):
    """conv4 .. conv6_1 (shared with FlowNetC / FlowNetSD); c3 is the 1/8-resolution 256-channel map inside cat[3]."""
    t = p.conv(L["conv4"], c3, p.act(B, c3.h // 2, c3.w // 2, 512))
    c4 = p.conv(L["conv4_1"], t, cat[4].view(0, 512))
    t = p.conv(L["conv5"], c4, p.act(B, c4.h // 2, c4.w // 2, 512))
    c5 = p.conv(L["conv5_1"], t, cat[5].view(0, 512))
    t = p.conv(L["conv6"], c5, p.act(B, c5.h // 2, c5.w // 2, 1024))
    return p.conv(L["conv6_1"], t, p.act(B, t.h, t.w, 1024))


def build(p, L, x, B, tag=""):
    """x: split-fp16 NHWC input view (6 -> 8 or 12 -> 16 channels).  Returns fp32 flow2 [B,2,H/4,W/4]."""
    H, W = x.h, x.w
    cat = sm.alloc_concats(p, B, H, W)
    t = p.conv(L["conv1"], x, p.act(B, H // 2, W // 2, 64))
    c2 = p.conv(L["conv2"], t, cat[2].view(0, 128))
    t = p.conv(L["conv3"], c2, p.act(B, H // 8, W // 8, 256))
    c3 = p.conv(L["conv3_1"], t, cat[3].view(0, 256))
    c6 = encoder_tail(p, L, cat, c3, B)
    return sm.decoder(p, L, cat, c6, B, inter=False, tag=tag)
