"""FlowNetFusion (lib/flownet/networks/FlowNetFusion.py:11-67): full-resolution 3-level U-net fusing the two candidates."""
import torch.nn as nn

from . import submodules as sm


class FlowNetFusion(nn.Module):
    def __init__(self, args=None, batchNorm=True):
        super().__init__()
        bn = self.batchNorm = batchNorm
        self.conv0, self.conv1, self.conv1_1 = sm.conv(bn, 11, 64), sm.conv(bn, 64, 64, stride=2), sm.conv(bn, 64, 128)
        self.conv2, self.conv2_1 = sm.conv(bn, 128, 128, stride=2), sm.conv(bn, 128, 128)
        self.deconv1, self.deconv0 = sm.deconv(128, 32), sm.deconv(162, 16)
        self.inter_conv1, self.inter_conv0 = sm.i_conv(bn, 162, 32), sm.i_conv(bn, 82, 16)
        self.predict_flow2, self.predict_flow1, self.predict_flow0 = sm.predict_flow(128), sm.predict_flow(32), sm.predict_flow(16)
        self.upsampled_flow2_to_1 = nn.ConvTranspose2d(2, 2, 4, 2, 1)
        self.upsampled_flow1_to_0 = nn.ConvTranspose2d(2, 2, 4, 2, 1)
        sm.flow_init(self)


def build(p, L, x, B, out_f32):
    """x: split-fp16 NHWC [B,H,W,16] (11 valid channels).  Writes the fused flow into out_f32 [B,2,H,W]."""
    H, W = x.h, x.w
    cat0 = p.act(B, H, W, 64 + 16 + 2, zero=True)
    cat1 = p.act(B, H // 2, W // 2, 128 + 32 + 2, zero=True)
    c0 = p.conv(L["conv0"], x, cat0.view(0, 64))
    t = p.conv(L["conv1"], c0, p.act(B, H // 2, W // 2, 64))
    c1 = p.conv(L["conv1_1"], t, cat1.view(0, 128))
    t = p.conv(L["conv2"], c1, p.act(B, H // 4, W // 4, 128))
    c2 = p.conv(L["conv2_1"], t, p.act(B, H // 4, W // 4, 128))
    f = p.flow_head(L["predict_flow2"], c2)
    p.flow_up(L["upsampled_flow2_to_1"], f, cat1.view(160, 2))
    p.conv(L["deconv1"], c2, cat1.view(128, 32))
    t = p.conv(L["inter_conv1"], cat1, p.act(B, cat1.h, cat1.w, 32))
    f = p.flow_head(L["predict_flow1"], t)
    p.flow_up(L["upsampled_flow1_to_0"], f, cat0.view(80, 2))
    p.conv(L["deconv0"], cat1, cat0.view(64, 16))
    t = p.conv(L["inter_conv0"], cat0, p.act(B, H, W, 16))
    return p.flow_head(L["predict_flow0"], t, out=out_f32, name="fusion")
