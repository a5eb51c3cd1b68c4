"""Layer builders with the reference's names and signatures (lib/flownet/networks/submodules.py:7-38).  The
returned modules are parameter containers (state_dict-compatible with reference checkpoints); the arithmetic runs
through engine.ConvLayer (bias or folded BatchNorm + LeakyReLU(0.1) fused into the conv epilogue)."""
import torch.nn as nn

from ... import _lib, engine


def conv(batchNorm, in_planes, out_planes, kernel_size=3, stride=1):
    layers = [nn.Conv2d(in_planes, out_planes, kernel_size, stride, (kernel_size - 1) // 2, bias=not batchNorm)]
    if batchNorm:
        layers.append(nn.BatchNorm2d(out_planes))
    return nn.Sequential(*layers, nn.LeakyReLU(0.1, inplace=True))


def i_conv(batchNorm, in_planes, out_planes, kernel_size=3, stride=1, bias=True):
    layers = [nn.Conv2d(in_planes, out_planes, kernel_size, stride, (kernel_size - 1) // 2, bias=bias)]
    if batchNorm:
        layers.append(nn.BatchNorm2d(out_planes))
    return nn.Sequential(*layers)


def predict_flow(in_planes):
    return nn.Conv2d(in_planes, 2, kernel_size=3, stride=1, padding=1, bias=True)


def deconv(in_planes, out_planes):
    return nn.Sequential(nn.ConvTranspose2d(in_planes, out_planes, 4, 2, 1, bias=True), nn.LeakyReLU(0.1, inplace=True))


def flow_init(module):
    """xavier-uniform weights, U(0,1) biases -- every FlowNet class of the reference ends its constructor with this
    (e.g. FlowNetS.py:47-56)."""
    for m in module.modules():
        if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
            if m.bias is not None:
                nn.init.uniform_(m.bias)
            nn.init.xavier_uniform_(m.weight)


def pack(mod, device, keep_host=False):
    """Sequential[(de)conv, (BN), (LeakyReLU)] or a bare conv -> engine.ConvLayer."""
    if isinstance(mod, nn.Sequential):
        mods = list(mod)
        bn = next((m for m in mods if isinstance(m, nn.BatchNorm2d)), None)
        act = _lib.ACT_LEAKY if any(isinstance(m, nn.LeakyReLU) for m in mods) else _lib.ACT_NONE
        return engine.conv_from_modules(mods[0], bn, act, device=device, keep_host=keep_host)
    return engine.conv_from_modules(mod, None, _lib.ACT_NONE, device=device, keep_host=keep_host)


def pack_all(net, device, keep_host=False):
    """Every direct child that is a conv-like block -> {attribute name: ConvLayer}; the two-channel flow heads and the
    2->2 flow up-samplers get their own packed forms (engine.FlowHead / engine.FlowUp)."""
    out = {}
    for name, child in net.named_children():
        if name.startswith("predict_flow"):
            out[name] = engine.FlowHead(child, device=device, keep_host=keep_host)
        elif name.startswith("upsampled_flow"):
            out[name] = engine.FlowUp(child, device=device)
        elif isinstance(child, (nn.Conv2d, nn.ConvTranspose2d)) or (isinstance(child, nn.Sequential) and len(child) and isinstance(child[0], (nn.Conv2d, nn.ConvTranspose2d))):
            out[name] = pack(child, device, keep_host)
    return out


# ---- the refinement ladder shared by FlowNetS / FlowNetC / FlowNetSD (FlowNetS.py:69-90, FlowNetSD.py:75-101) ----
SKIP = {5: (512, 512), 4: (512, 256), 3: (256, 128), 2: (128, 64)}     # level -> (skip channels, deconv channels)


def alloc_concats(p, B, H, W):
    """Channel-concatenated buffers [skip | deconv | flow_up(2 -> 8)] at 1/32 .. 1/4 resolution; producers write slices."""
    return {lvl: p.act(B, H >> lvl, W >> lvl, s + d + 2, zero=True) for lvl, (s, d) in SKIP.items()}


def decoder(p, L, cat, c6, B, inter, tag=""):
    """Adds the ladder to plan `p`; returns the fp32 NCHW flow2 tensor [B,2,H/4,W/4]."""
    # at every level the flow head (an 18-output 1x1 convolution on a handful of tiles + two tiny kernels) and the deconvolution read
    # the same concat buffer and write disjoint slices of the next one: two branches of the plan (parallel paths of the CUDA graph)
    with p.branch() as br:
        f = p.flow_head(L["predict_flow6"], c6)
        p.flow_up(L["upsampled_flow6_to_5"], f, cat[5].view(1024, 2))
    p.conv(L["deconv5"], c6, cat[5].view(512, 512))
    br.join()
    for lvl in (5, 4, 3):
        src = cat[lvl]
        s, d = SKIP[lvl - 1]
        with p.branch() as br:
            head_in = p.conv(L["inter_conv%d" % lvl], src, p.act(B, src.h, src.w, L["inter_conv%d" % lvl].cout)) if inter else src
            f = p.flow_head(L["predict_flow%d" % lvl], head_in)
            p.flow_up(L["upsampled_flow%d_to_%d" % (lvl, lvl - 1)], f, cat[lvl - 1].view(s + d, 2))
        p.conv(L["deconv%d" % (lvl - 1)], src, cat[lvl - 1].view(s, d))
        br.join()
    src = cat[2]
    head_in = p.conv(L["inter_conv2"], src, p.act(B, src.h, src.w, L["inter_conv2"].cout)) if inter else src
    return p.flow_head(L["predict_flow2"], head_in, name=tag + "flow2")


def add_decoder_modules(net, bn, up_bias, inter):
    net.deconv5, net.deconv4, net.deconv3, net.deconv2 = deconv(1024, 512), deconv(1026, 256), deconv(770, 128), deconv(386, 64)
    if inter:
        net.inter_conv5, net.inter_conv4 = i_conv(bn, 1026, 512), i_conv(bn, 770, 256)
        net.inter_conv3, net.inter_conv2 = i_conv(bn, 386, 128), i_conv(bn, 194, 64)
    widths = (1024, 512, 256, 128, 64) if inter else (1024, 1026, 770, 386, 194)
    for lvl, c in zip((6, 5, 4, 3, 2), widths):
        setattr(net, "predict_flow%d" % lvl, predict_flow(c))
    for lvl in (6, 5, 4, 3):
        setattr(net, "upsampled_flow%d_to_%d" % (lvl, lvl - 1), nn.ConvTranspose2d(2, 2, 4, 2, 1, bias=up_bias))
