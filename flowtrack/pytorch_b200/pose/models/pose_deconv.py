"""PoseResNet ("DeconvResnet") drop-in: reference interface, B200 kernels underneath.

Interface kept from the reference (lib/pose/models/pose_deconv.py:12-46,165-179; resnet.py:15-55;
blocks.py:83-120): `deconv(backbone, num_classes, pretrained)` factory, `forward(x[B,3,H,W]) ->
[B,num_classes,H/4,W/4]`, and the state_dict key layout (`conv1.weight`, `bn1.*`,
`layer{1..4}.{i}.conv{1,2,3}.weight`, `...downsample.{0,1}.*`, `deconv.{0,3,6}.weight`, `deconv.{1,4,7}.*`,
`heatmap.{weight,bias}`), so reference checkpoints load with strict=True.

The nn.Conv2d / BatchNorm2d children are parameter containers only; forward() never calls them.  On the
first forward (and after load_state_dict / .to()) the weights are folded and packed (engine.py) and the
forward becomes a fixed list of launches of the tcgen05 implicit-GEMM kernel:
  stem 7x7/2 (+BN+ReLU) -> max-pool -> 16/33/50 bottlenecks (1x1, 3x3, 1x1 with BN folded, the residual
  add and ReLU fused into the third conv's epilogue; projection shortcut as its own conv) -> three
  k4 s2 deconvs as 4-phase 2x2 convs (+BN+ReLU) -> 1x1 head writing fp32 NCHW heat-maps.
"""
import os

import torch
import torch.nn as nn

from ... import _lib, engine

_STAGES = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3), 152: (3, 8, 36, 3)}


class _Bottleneck(nn.Module):
    """Parameter container with the reference's attribute names (blocks.py:83-103)."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1):
        super().__init__()
        wide = planes * self.expansion
        self.conv1, self.bn1 = nn.Conv2d(inplanes, planes, 1, bias=False), nn.BatchNorm2d(planes)
        self.conv2, self.bn2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False), nn.BatchNorm2d(planes)
        self.conv3, self.bn3 = nn.Conv2d(planes, wide, 1, bias=False), nn.BatchNorm2d(wide)
        self.downsample = nn.Sequential()
        if stride != 1 or inplanes != wide:
            self.downsample = nn.Sequential(nn.Conv2d(inplanes, wide, 1, stride, bias=False), nn.BatchNorm2d(wide))
        self.stride = stride


class DeconvResnet(nn.Module):
    def __init__(self, depth=50, num_classes=17, precision="split"):
        super().__init__()
        if depth not in _STAGES:
            raise ValueError("the B200 hot path implements the Bottleneck trunks (resnet50/101/152); got resnet%d" % depth)
        self.depth, self.num_classes = depth, num_classes
        self.precision = precision          # "split" (fp32-grade, default) | "fast" (fp16 operands)
        self.use_cuda_graph = True
        self.conv1, self.bn1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False), nn.BatchNorm2d(64)
        inplanes = 64
        for i, (planes, count) in enumerate(zip((64, 128, 256, 512), _STAGES[depth])):
            blocks = []
            for b in range(count):
                blocks.append(_Bottleneck(inplanes, planes, 2 if (b == 0 and i > 0) else 1))
                inplanes = planes * 4
            setattr(self, "layer%d" % (i + 1), nn.Sequential(*blocks))
        mods = []
        for i in range(3):
            mods += [nn.ConvTranspose2d(inplanes if i == 0 else 256, 256, 4, 2, 1, bias=False), nn.BatchNorm2d(256), nn.ReLU(inplace=True)]
        self.deconv = nn.Sequential(*mods)
        self.heatmap = nn.Conv2d(256, num_classes, 1)
        self._layers = None     # packed ConvLayers (device)
        self._plans = {}        # (B,H,W,precision,impl) -> (plan, input act, output tensor)
        self._impl = _lib.IMPL_TCGEN05
        self.eval()

    # ---- reference init (resnet.py:38-49, pose_deconv.py:48-63) -------------------------------------
    def init_weights(self, pretrained=""):
        if os.path.isfile(pretrained):
            self.load_state_dict(torch.load(pretrained), strict=False)
            return
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def init_net(self, pretrained=""):
        self.init_weights(pretrained)
        for m in self.deconv.modules():
            if isinstance(m, nn.ConvTranspose2d):
                nn.init.normal_(m.weight, std=0.001)
        nn.init.normal_(self.heatmap.weight, std=0.001)
        nn.init.constant_(self.heatmap.bias, 0)
        self.refresh()

    # ---- packed-weight cache invalidation -----------------------------------------------------------
    def refresh(self):
        """Drop packed weights and plans; call after editing parameters in place."""
        self._layers, self._plans = None, {}

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        self.refresh()
        return r

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self.refresh()
        return r

    def train(self, mode=True):
        if mode:
            raise RuntimeError("flowtrack.pytorch_b200 implements the inference path only (eval-mode BatchNorm is folded into the convolutions)")
        return super().train(False)

    # ---- packing ------------------------------------------------------------------------------------
    def _pack(self, device, keep_host=False):
        R, N = _lib.ACT_RELU, _lib.ACT_NONE
        mk = lambda conv, bn, act, cin_p=None, residual=False: engine.conv_from_modules(conv, bn, act, cin_p=cin_p, device=device, keep_host=keep_host,
                                                                                        residual=residual)
        # stem: space-to-depth form (4 k-blocks per tile instead of 7, half the packed input bytes); FT_STEM_S2D=0 keeps the 7x7 form
        self._stem_s2d = os.environ.get("FT_STEM_S2D", "1") != "0"
        stem = engine.stem_s2d_from_modules(self.conv1, self.bn1, R, device=device, keep_host=keep_host) if self._stem_s2d else mk(self.conv1, self.bn1, R, 8)
        L = {"stem": stem, "blocks": []}
        for li in range(1, 5):
            for blk in getattr(self, "layer%d" % li):
                L["blocks"].append(dict(
                    c1=mk(blk.conv1, blk.bn1, R), c2=mk(blk.conv2, blk.bn2, R), c3=mk(blk.conv3, blk.bn3, R, residual=True),  # residual add, then ReLU
                    down=mk(blk.downsample[0], blk.downsample[1], N) if len(blk.downsample) else None))
        L["deconv"] = [mk(self.deconv[3 * i], self.deconv[3 * i + 1], R) for i in range(3)]
        L["head"] = mk(self.heatmap, None, N)
        return L

    def _build_plan(self, B, H, W, device, precision, impl):
        if self._layers is None:
            self._layers = self._pack(device)
        L = self._layers
        p = engine.Plan(device, precision, impl)
        src = p.f32(B, 3, H, W)                                 # static fp32 NCHW input (graph-capturable)
        if L["stem"].cin_p == 16:                               # space-to-depth stem: [B, H/2, W/2, 16], rows zero-padded for FT_K_ROW
            x = p.act(B, H // 2, W // 2, 16, xpad=L["stem"].in_xpad)
            p.call("pack_image_s2d", _lib.lib().ft_pack_image_s2d, src.data_ptr(), B, H, W, x.struct(), x.xl, x.xr)
            p.taps["input_s2d"] = x                             # a caller may fill this itself (ft_crop_affine_s2d) and enqueue(start=1)
            h, w = L["stem"].out_hw(H // 2, W // 2)
        else:
            x = p.act(B, H, W, 8, xpad=L["stem"].in_xpad)       # packed input, 3 -> 8 channels, rows zero-padded for the FT_K_ROW stem
            xs, st = x.struct(), src.stride()
            p.call("pack_image", _lib.lib().ft_pack_image, src.data_ptr(), st[0], st[1], st[2], st[3], 3, 1.0, xs, x.xl, x.xr)
            h, w = L["stem"].out_hw(H, W)
        y = p.conv(L["stem"], x, p.act(B, h, w, 64), name="stem")
        h, w = (h + 1) // 2, (w + 1) // 2
        pooled = p.act(B, h, w, 64)
        p.call("maxpool", _lib.lib().ft_maxpool3x3s2, y.struct(), pooled.struct())
        p.taps["pool"] = pooled
        cur = pooled
        for i, blk in enumerate(L["blocks"]):
            h2, w2 = blk["c2"].out_hw(cur.h, cur.w)
            br = None
            if blk["down"] is not None:           # the projection shortcut only needs the block input: a branch next to conv1 -> conv2
                with p.branch() as br:
                    res = p.conv(blk["down"], cur, p.act(B, h2, w2, blk["down"].cout))
            else:
                res = cur
            t1 = p.conv(blk["c1"], cur, p.act(B, cur.h, cur.w, blk["c1"].cout))
            t2 = p.conv(blk["c2"], t1, p.act(B, h2, w2, blk["c2"].cout))
            if br is not None:
                br.join()
            cur = p.conv(blk["c3"], t2, p.act(B, h2, w2, blk["c3"].cout), res=res, name="block%d" % i)
        for i, dl in enumerate(L["deconv"]):
            cur = p.conv(dl, cur, p.act(B, 2 * cur.h, 2 * cur.w, 256), name="deconv%d" % i)
        out = p.f32(B, self.num_classes, cur.h, cur.w)
        p.conv(L["head"], cur, None, out_f32=out, name="heatmap")
        return p, src, out

    def plan_for(self, B, H, W, device):
        """(plan, static fp32 NCHW input [B,3,H,W], static fp32 NCHW heat-maps) for this batch shape; built once and
        cached.  tracking/pipeline.py captures several plans into one CUDA graph through this."""
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("DeconvResnet.forward needs a CUDA tensor: this package has no CPU path")
        if H % 32 or W % 32:
            raise ValueError("expected [B,3,H,W] with H, W multiples of 32, got %s" % ((B, 3, H, W),))
        prec = _lib.PREC_FAST if self.precision == "fast" else _lib.PREC_SPLIT
        key = (B, H, W, prec, self._impl, device.index if device.index is not None else torch.cuda.current_device())
        if key not in self._plans:
            self._plans[key] = self._build_plan(B, H, W, device, prec, self._impl)
            if os.environ.get("FT_AUTOTUNE", "1") != "0":
                self._plans[key][0].autotune()       # per-layer tile shape by measurement (engine.Plan.autotune)
        return self._plans[key]

    def _plan(self, x):
        if not x.is_cuda:
            raise RuntimeError("DeconvResnet.forward needs a CUDA tensor: this package has no CPU path")
        B, C, H, W = x.shape
        if C != 3:
            raise ValueError("expected [B,3,H,W] with H, W multiples of 32, got %s" % (tuple(x.shape),))
        return self.plan_for(B, H, W, x.device)

    @torch.no_grad()
    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("DeconvResnet.forward needs a CUDA tensor: this package has no CPU path")
        with torch.cuda.device(x.device):                             # launches follow the tensor's device, not the caller's current one
            plan, src, out = self._plan(x)
            src.copy_(x)
            if self.use_cuda_graph and plan.graph is None and not torch.cuda.is_current_stream_capturing():
                plan.capture()
            plan.run()
            return out.clone()


def deconv(backbone, num_classes, pretrained):
    """Reference factory (pose_deconv.py:165-179): deconv('resnet50', 17, False)."""
    if not backbone.startswith("resnet"):
        raise ValueError("only the resnet backbones are on the B200 hot path (got %r)" % backbone)
    model = DeconvResnet(int(backbone[len("resnet"):]), num_classes)
    model.init_net(os.path.join("data", "pretrained", "%s.pth" % backbone) if pretrained else "")
    return model
