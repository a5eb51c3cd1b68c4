"""oracle/flow_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Plain-torch fp32 CPU restatement of the reference FlowNet2 family:

  builders      lib/flownet/networks/submodules.py:7-38  (conv / i_conv / predict_flow / deconv)
  FlowNetS      lib/flownet/networks/FlowNetS.py:15-95
  FlowNetC      lib/flownet/networks/FlowNetC.py:13-128
  FlowNetSD     lib/flownet/networks/FlowNetSD.py:11-106
  FlowNetFusion lib/flownet/networks/FlowNetFusion.py:11-67
  wrappers      lib/flownet/model/models.py:19-498 (FlowNet2, FlowNet2C/S/SD/CS/CSS)

The three native ops come from oracle/ops_oracle.py (CPU restatements; the reference has no CPU
implementation of them).  state_dict keys equal the reference's; tests/golden/make_golden.py loads
these weights into the real reference modules (imported through a torch.utils.ffi stub layer) and
stores the outputs as fixtures, which pins this restatement.

Not restated: `.register_hook(save_grad(...))` bookkeeping (models.py:151-176, training only) and
the fp16 `tofp32()/tofp16()` wrappers (args.fp16 is False on the path).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops_oracle as ops


def _conv(bn, cin, cout, k=3, stride=1):
    layers = [nn.Conv2d(cin, cout, k, stride=stride, padding=(k - 1) // 2, bias=not bn)]
    if bn:
        layers.append(nn.BatchNorm2d(cout))
    layers.append(nn.LeakyReLU(0.1, inplace=True))
    return nn.Sequential(*layers)


def _iconv(bn, cin, cout):
    layers = [nn.Conv2d(cin, cout, 3, padding=1, bias=True)]
    if bn:
        layers.append(nn.BatchNorm2d(cout))
    return nn.Sequential(*layers)


def _deconv(cin, cout):
    return nn.Sequential(nn.ConvTranspose2d(cin, cout, 4, stride=2, padding=1, bias=True), nn.LeakyReLU(0.1, inplace=True))


def _pred(cin):
    return nn.Conv2d(cin, 2, 3, padding=1, bias=True)


def _flow_init(mod):
    """xavier-uniform weights, U(0,1) biases (FlowNetS.py:47-56 and every sibling class)."""
    for m in mod.modules():
        if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
            if m.bias is not None:
                nn.init.uniform_(m.bias)
            nn.init.xavier_uniform_(m.weight)


def _decoder(self, feats, conv6, inter):
    """Shared refinement ladder (FlowNetS.py:69-90, FlowNetC.py:103-123, FlowNetSD.py:75-101).
    feats = (skip2, skip3, skip4, skip5) encoder maps; returns flow2..flow6."""
    skip2, skip3, skip4, skip5 = feats
    flow6 = self.predict_flow6(conv6)
    up = self.upsampled_flow6_to_5(flow6)
    cat = torch.cat((skip5, self.deconv5(conv6), up), 1)
    flows = [flow6]
    for lvl, skip in ((5, skip4), (4, skip3), (3, skip2)):
        pf = getattr(self, "predict_flow%d" % lvl)
        flow = pf(getattr(self, "inter_conv%d" % lvl)(cat)) if inter else pf(cat)
        flows.append(flow)
        up = getattr(self, "upsampled_flow%d_to_%d" % (lvl, lvl - 1))(flow)
        cat = torch.cat((skip, getattr(self, "deconv%d" % (lvl - 1))(cat), up), 1)
    flow2 = self.predict_flow2(self.inter_conv2(cat)) if inter else self.predict_flow2(cat)
    flows.append(flow2)
    return tuple(reversed(flows))  # flow2, flow3, flow4, flow5, flow6


def _add_decoder(self, bn, up_bias, inter):
    self.deconv5, self.deconv4, self.deconv3, self.deconv2 = _deconv(1024, 512), _deconv(1026, 256), _deconv(770, 128), _deconv(386, 64)
    if inter:
        self.inter_conv5, self.inter_conv4 = _iconv(bn, 1026, 512), _iconv(bn, 770, 256)
        self.inter_conv3, self.inter_conv2 = _iconv(bn, 386, 128), _iconv(bn, 194, 64)
        pin = (1024, 512, 256, 128, 64)
    else:
        pin = (1024, 1026, 770, 386, 194)
    for lvl, c in zip((6, 5, 4, 3, 2), pin):
        setattr(self, "predict_flow%d" % lvl, _pred(c))
    for lvl in (6, 5, 4, 3):
        setattr(self, "upsampled_flow%d_to_%d" % (lvl, lvl - 1), nn.ConvTranspose2d(2, 2, 4, 2, 1, bias=up_bias))


class FlowNetS(nn.Module):
    def __init__(self, args=None, input_channels=12, batchNorm=True):
        super().__init__()
        bn = self.batchNorm = batchNorm
        self.conv1, self.conv2, self.conv3 = _conv(bn, input_channels, 64, 7, 2), _conv(bn, 64, 128, 5, 2), _conv(bn, 128, 256, 5, 2)
        self.conv3_1, self.conv4, self.conv4_1 = _conv(bn, 256, 256), _conv(bn, 256, 512, 3, 2), _conv(bn, 512, 512)
        self.conv5, self.conv5_1 = _conv(bn, 512, 512, 3, 2), _conv(bn, 512, 512)
        self.conv6, self.conv6_1 = _conv(bn, 512, 1024, 3, 2), _conv(bn, 1024, 1024)
        _add_decoder(self, bn, up_bias=False, inter=False)
        _flow_init(self)
        self.upsample1 = nn.Upsample(scale_factor=4, mode="bilinear")

    def features(self, x):
        c2 = self.conv2(self.conv1(x))
        c3 = self.conv3_1(self.conv3(c2))
        c4 = self.conv4_1(self.conv4(c3))
        c5 = self.conv5_1(self.conv5(c4))
        c6 = self.conv6_1(self.conv6(c5))
        return (c2, c3, c4, c5), c6

    def forward(self, x):
        feats, c6 = self.features(x)
        flows = _decoder(self, feats, c6, inter=False)
        return flows if self.training else (flows[0],)


class FlowNetC(nn.Module):
    def __init__(self, args=None, batchNorm=True, div_flow=20):
        super().__init__()
        bn = self.batchNorm = batchNorm
        self.div_flow = div_flow
        self.conv1, self.conv2, self.conv3 = _conv(bn, 3, 64, 7, 2), _conv(bn, 64, 128, 5, 2), _conv(bn, 128, 256, 5, 2)
        self.conv_redir = _conv(bn, 256, 32, 1, 1)
        self.conv3_1, self.conv4, self.conv4_1 = _conv(bn, 473, 256), _conv(bn, 256, 512, 3, 2), _conv(bn, 512, 512)
        self.conv5, self.conv5_1 = _conv(bn, 512, 512, 3, 2), _conv(bn, 512, 512)
        self.conv6, self.conv6_1 = _conv(bn, 512, 1024, 3, 2), _conv(bn, 1024, 1024)
        _add_decoder(self, bn, up_bias=True, inter=False)
        _flow_init(self)
        self.upsample1 = nn.Upsample(scale_factor=4, mode="bilinear")

    def features(self, x):
        a2 = self.conv2(self.conv1(x[:, 0:3]))
        a3 = self.conv3(a2)
        b3 = self.conv3(self.conv2(self.conv1(x[:, 3:])))
        corr = F.leaky_relu(ops.correlation_torch(a3, b3, 20, 1, 20, 1, 2), 0.1)   # FlowNetC.py:31,86-87
        c3 = self.conv3_1(torch.cat((self.conv_redir(a3), corr), 1))
        c4 = self.conv4_1(self.conv4(c3))
        c5 = self.conv5_1(self.conv5(c4))
        c6 = self.conv6_1(self.conv6(c5))
        return (a2, c3, c4, c5), c6

    def forward(self, x):
        feats, c6 = self.features(x)
        flows = _decoder(self, feats, c6, inter=False)
        return flows if self.training else (flows[0],)


class FlowNetSD(nn.Module):
    def __init__(self, args=None, batchNorm=True):
        super().__init__()
        bn = self.batchNorm = batchNorm
        self.conv0, self.conv1, self.conv1_1 = _conv(bn, 6, 64), _conv(bn, 64, 64, 3, 2), _conv(bn, 64, 128)
        self.conv2, self.conv2_1 = _conv(bn, 128, 128, 3, 2), _conv(bn, 128, 128)
        self.conv3, self.conv3_1 = _conv(bn, 128, 256, 3, 2), _conv(bn, 256, 256)
        self.conv4, self.conv4_1 = _conv(bn, 256, 512, 3, 2), _conv(bn, 512, 512)
        self.conv5, self.conv5_1 = _conv(bn, 512, 512, 3, 2), _conv(bn, 512, 512)
        self.conv6, self.conv6_1 = _conv(bn, 512, 1024, 3, 2), _conv(bn, 1024, 1024)
        _add_decoder(self, bn, up_bias=True, inter=True)
        _flow_init(self)
        self.upsample1 = nn.Upsample(scale_factor=4, mode="bilinear")

    def forward(self, x):
        c1 = self.conv1_1(self.conv1(self.conv0(x)))
        c2 = self.conv2_1(self.conv2(c1))
        c3 = self.conv3_1(self.conv3(c2))
        c4 = self.conv4_1(self.conv4(c3))
        c5 = self.conv5_1(self.conv5(c4))
        c6 = self.conv6_1(self.conv6(c5))
        flows = _decoder(self, (c2, c3, c4, c5), c6, inter=True)
        return flows if self.training else (flows[0],)


class FlowNetFusion(nn.Module):
    def __init__(self, args=None, batchNorm=True):
        super().__init__()
        bn = self.batchNorm = batchNorm
        self.conv0, self.conv1, self.conv1_1 = _conv(bn, 11, 64), _conv(bn, 64, 64, 3, 2), _conv(bn, 64, 128)
        self.conv2, self.conv2_1 = _conv(bn, 128, 128, 3, 2), _conv(bn, 128, 128)
        self.deconv1, self.deconv0 = _deconv(128, 32), _deconv(162, 16)
        self.inter_conv1, self.inter_conv0 = _iconv(bn, 162, 32), _iconv(bn, 82, 16)
        self.predict_flow2, self.predict_flow1, self.predict_flow0 = _pred(128), _pred(32), _pred(16)
        self.upsampled_flow2_to_1 = nn.ConvTranspose2d(2, 2, 4, 2, 1)
        self.upsampled_flow1_to_0 = nn.ConvTranspose2d(2, 2, 4, 2, 1)
        _flow_init(self)

    def forward(self, x):
        c0 = self.conv0(x)
        c1 = self.conv1_1(self.conv1(c0))
        c2 = self.conv2_1(self.conv2(c1))
        flow2 = self.predict_flow2(c2)
        cat1 = torch.cat((c1, self.deconv1(c2), self.upsampled_flow2_to_1(flow2)), 1)
        flow1 = self.predict_flow1(self.inter_conv1(cat1))
        cat0 = torch.cat((c0, self.deconv0(cat1), self.upsampled_flow1_to_0(flow1)), 1)
        return self.predict_flow0(self.inter_conv0(cat0))


def normalise_pair(inputs, rgb_max):
    """models.py:112-119 -- per-sample, per-colour mean over both frames and all pixels; /rgb_max; stack on C."""
    mean = inputs.contiguous().view(inputs.size()[:2] + (-1,)).mean(dim=-1).view(inputs.size()[:2] + (1, 1, 1))
    x = (inputs - mean) / rgb_max
    return torch.cat((x[:, :, 0], x[:, :, 1]), dim=1)


def _up4(t, mode):
    return F.interpolate(t, scale_factor=4, mode=mode, align_corners=False if mode == "bilinear" else None)


def _warp_block(x, flow, div_flow):
    """models.py:126-131: warp img1 by flow, brightness error, its channel norm, 12-channel concat."""
    warped = ops.resample2d_torch(x[:, 3:], flow)
    norm = ops.channelnorm_torch(x[:, :3] - warped)
    return torch.cat((x, warped, flow / div_flow, norm), dim=1)


class _Args:
    rgb_max = 255.0
    fp16 = False
    grads = {}


class FlowNet2S(FlowNetS):
    def __init__(self, args=None, batchNorm=False, div_flow=20):
        super().__init__(args, input_channels=6, batchNorm=batchNorm)
        self.rgb_max = (args or _Args).rgb_max
        self.div_flow = div_flow

    def forward(self, inputs):
        flows = FlowNetS.forward(self, normalise_pair(inputs, self.rgb_max))
        return flows if self.training else _up4(flows[0] * self.div_flow, "bilinear")


class FlowNet2C(FlowNetC):
    def __init__(self, args=None, batchNorm=False, div_flow=20):
        super().__init__(args, batchNorm=batchNorm, div_flow=20)
        self.rgb_max = (args or _Args).rgb_max

    def forward(self, inputs):
        flows = FlowNetC.forward(self, normalise_pair(inputs, self.rgb_max))
        return flows if self.training else _up4(flows[0] * self.div_flow, "bilinear")


class FlowNet2SD(FlowNetSD):
    def __init__(self, args=None, batchNorm=False, div_flow=20):
        super().__init__(args, batchNorm=batchNorm)
        self.rgb_max = (args or _Args).rgb_max
        self.div_flow = div_flow

    def forward(self, inputs):
        flows = FlowNetSD.forward(self, normalise_pair(inputs, self.rgb_max))
        return flows if self.training else _up4(flows[0] * self.div_flow, "bilinear")


class FlowNet2CS(nn.Module):
    def __init__(self, args=None, batchNorm=False, div_flow=20.0):
        super().__init__()
        self.batchNorm, self.div_flow, self.rgb_max = batchNorm, div_flow, (args or _Args).rgb_max
        self.flownetc = FlowNetC(args, batchNorm=batchNorm)
        self.flownets_1 = FlowNetS(args, batchNorm=batchNorm)
        _flow_init(self)

    def forward(self, inputs):
        x = normalise_pair(inputs, self.rgb_max)
        flow = _up4(self.flownetc(x)[0] * self.div_flow, "bilinear")
        return _up4(self.flownets_1(_warp_block(x, flow, self.div_flow))[0] * self.div_flow, "bilinear")


class FlowNet2CSS(nn.Module):
    def __init__(self, args=None, batchNorm=False, div_flow=20.0):
        super().__init__()
        self.batchNorm, self.div_flow, self.rgb_max = batchNorm, div_flow, (args or _Args).rgb_max
        self.flownetc = FlowNetC(args, batchNorm=batchNorm)
        self.flownets_1 = FlowNetS(args, batchNorm=batchNorm)
        self.flownets_2 = FlowNetS(args, batchNorm=batchNorm)
        _flow_init(self)

    def forward(self, inputs):
        x = normalise_pair(inputs, self.rgb_max)
        flow = _up4(self.flownetc(x)[0] * self.div_flow, "bilinear")
        flow = _up4(self.flownets_1(_warp_block(x, flow, self.div_flow))[0] * self.div_flow, "bilinear")
        return _up4(self.flownets_2(_warp_block(x, flow, self.div_flow))[0] * self.div_flow, "nearest")


class FlowNet2(nn.Module):
    def __init__(self, args=None, batchNorm=False, div_flow=20.0):
        super().__init__()
        self.batchNorm, self.div_flow, self.rgb_max = batchNorm, div_flow, (args or _Args).rgb_max
        self.flownetc = FlowNetC(args, batchNorm=batchNorm)
        self.flownets_1 = FlowNetS(args, batchNorm=batchNorm)
        self.flownets_2 = FlowNetS(args, batchNorm=batchNorm)
        self.flownets_d = FlowNetSD(args, batchNorm=batchNorm)
        self.flownetfusion = FlowNetFusion(args, batchNorm=batchNorm)
        _flow_init(self)

    def forward(self, inputs):
        x = normalise_pair(inputs, self.rgb_max)
        flow = _up4(self.flownetc(x)[0] * self.div_flow, "bilinear")                                      # :121-122
        flow = _up4(self.flownets_1(_warp_block(x, flow, self.div_flow))[0] * self.div_flow, "bilinear")  # :126-135
        s2 = _up4(self.flownets_2(_warp_block(x, flow, self.div_flow))[0] * self.div_flow, "nearest")     # :138-147
        s2_norm = ops.channelnorm_torch(s2)
        s2_err = ops.channelnorm_torch(x[:, :3] - ops.resample2d_torch(x[:, 3:], s2))                     # :150-154
        sd = _up4(self.flownets_d(x)[0] / self.div_flow, "nearest")                                       # :159-160 (note '/')
        sd_norm = ops.channelnorm_torch(sd)
        sd_err = ops.channelnorm_torch(x[:, :3] - ops.resample2d_torch(x[:, 3:], sd))                     # :163-167
        cat = torch.cat((x[:, :3], sd, s2, sd_norm, s2_norm, sd_err, s2_err), dim=1)                      # :172
        return self.flownetfusion(cat)
