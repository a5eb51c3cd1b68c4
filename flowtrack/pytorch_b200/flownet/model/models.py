"""FlowNet2 family drop-ins: reference constructors / forward signatures (lib/flownet/model/models.py:19-498),
B200 kernels underneath.

    model = models.FlowNet2S(args)            # args.rgb_max (255.), args.fp16 (False), args.grads ({})
    flow  = model(inputs)                     # inputs [B,3,2,H,W] RGB 0..255 (CUDA fp32) -> [B,2,H,W] pixels (u, v)

state_dict keys equal the reference's (`conv1.0.weight`, `deconv5.0.bias`, `predict_flow6.weight`,
`flownetc.conv1.0.weight`, ...), so reference checkpoints (`ckpt['state_dict']`, tools.py:138-143) load strict.
Inference only: eval-mode outputs (the full-resolution flow); the training tuple (flow2..flow6) is out of scope.
args.fp16 must be False -- the path computes in split-fp16 (fp32-grade) or, with precision='fast', fp16 operands.
"""
import os

import torch
import torch.nn as nn

from ... import _lib, engine
from ..networks import FlowNetC, FlowNetFusion, FlowNetS, FlowNetSD
from ..networks import submodules as sm
from ..networks.channelnorm_package.modules.channelnorm import ChannelNorm
from ..networks.resample2d_package.modules.resample2d import Resample2d


class _Args:
    rgb_max, fp16, grads = 255.0, False, {}


class _FlowBase(nn.Module):
    """Plan cache + weight packing shared by every FlowNet2* wrapper."""

    def _init_base(self, args, div_flow):
        args = args if args is not None else _Args
        if getattr(args, "fp16", False):
            raise ValueError("args.fp16=True is not supported: use precision='fast' for the fp16-operand mode")
        self.rgb_max, self.div_flow, self.args = float(args.rgb_max), div_flow, args
        self.precision, self.use_cuda_graph, self._impl = "split", True, _lib.IMPL_TCGEN05
        self._layers, self._plans = None, {}

    def refresh(self):
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
            raise RuntimeError("flowtrack.pytorch_b200 implements the inference path only")
        return super().train(False)

    def _subnets(self):
        """name -> module holding conv children; single-net wrappers are their own sub-net."""
        return {"": self}

    def _pack(self, device, keep_host=False):
        return {name: sm.pack_all(mod, device, keep_host) for name, mod in self._subnets().items()}

    def _build(self, p, L, pair, B, H, W):
        raise NotImplementedError

    def plan_for(self, B, H, W, device, u8=False, tag=0):
        """(plan, static input, static fp32 flow [B,2,H,W]) for this batch shape; built once and cached.  The static input is the
        reference's fp32 [B,3,2,H,W] RGB pair, or with u8=True a uint8 [B,2,H,W,3] BGR buffer -- the cv2 frames the reference's
        callers hold (lib/tracking/net_utils.py:83-87), converted and normalised by the packing kernel.  `tag` lets a caller own
        several independent plans (own buffers) of one shape, e.g. the two lanes of tracking.tracker.TrackingPipeline."""
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("%s.forward needs a CUDA tensor: this package has no CPU path" % type(self).__name__)
        if H % 64 or W % 64:
            raise ValueError("expected [B,3,2,H,W] with H, W multiples of 64, got %s" % ((B, 3, 2, H, W),))
        prec = _lib.PREC_FAST if self.precision == "fast" else _lib.PREC_SPLIT
        key = (B, H, W, prec, self._impl, device.index if device.index is not None else torch.cuda.current_device(), bool(u8), tag)
        if key not in self._plans:
            if self._layers is None:
                self._layers = self._pack(device)
            p = engine.Plan(device, prec, self._impl)
            if u8:
                pair = torch.zeros(B, 2, H, W, 3, dtype=torch.uint8, device=device)
                p.keep.append(pair)
                p.pair_u8 = True
            else:
                pair = p.f32(B, 3, 2, H, W)
            out = self._build(p, self._layers, pair, B, H, W)
            if os.environ.get("FT_AUTOTUNE", "1") != "0":
                # per-layer tile shape by measurement (engine.Plan.autotune); a plan of the same shape (another lane) lends its choices
                twin = next((v[0].tuned for k, v in self._plans.items() if k[:-1] == key[:-1] and v[0].tuned), None)
                p.autotune(hint=twin)
            self._plans[key] = (p, pair, out)
        return self._plans[key]

    @torch.no_grad()
    def forward(self, inputs):
        if not inputs.is_cuda:
            raise RuntimeError("%s.forward needs a CUDA tensor: this package has no CPU path" % type(self).__name__)
        if inputs.dim() != 5 or inputs.shape[1] != 3 or inputs.shape[2] != 2 or inputs.shape[3] % 64 or inputs.shape[4] % 64:
            raise ValueError("expected [B,3,2,H,W] with H, W multiples of 64, got %s" % (tuple(inputs.shape),))
        B, _, _, H, W = inputs.shape
        with torch.cuda.device(inputs.device):                        # launches follow the tensor's device, not the caller's current one
            p, pair, out = self.plan_for(B, H, W, inputs.device)
            pair.copy_(inputs)
            if self.use_cuda_graph and p.graph is None and not torch.cuda.is_current_stream_capturing():
                p.capture()
            p.run()
            return out.clone()


# ---- plan fragments ---------------------------------------------------------------------------------
def _normalise(p, pair, B, H, W, want_f32, want_act, xpad=(0, 0)):
    """models.py:112-119: x = (pair - mean_bc)/rgb_max as [B,6,H,W] fp32 and/or split-fp16 NHWC (6 -> 8 channels).
    `xpad`: zero columns the consuming stem convolution wants around every row of the NHWC copy (ConvLayer.in_xpad)."""
    lib = _lib.lib()
    xf = p.f32(B, 6, H, W) if want_f32 else None
    xa = p.act(B, H, W, 6, xpad=xpad) if want_act else None
    if getattr(p, "pair_u8", False):          # uint8 BGR frames [B,2,H,W,3] (cv2 layout): flip to RGB, convert, normalise in one pass
        ws = torch.zeros(B * 3, dtype=torch.int64, device=p.device)
        p.keep.append(ws)
        p.call("pack_pair_u8", lib.ft_pack_pair_u8, pair.data_ptr(), B, H, W, 1, p.rgb_max, ws.data_ptr(),
               xf.data_ptr() if want_f32 else None, xa.struct() if want_act else None, xpad[0], xpad[1])
        return xf, xa
    ws = p.f32(lib.ft_pack_pair_ws_floats(B))
    p.call("pack_pair", lib.ft_pack_pair, pair.data_ptr(), B, H, W, p.rgb_max, ws.data_ptr(),
           xf.data_ptr() if want_f32 else None, xa.struct() if want_act else None, xpad[0], xpad[1])
    return xf, xa


def _up4(p, flow2, B, mul, mode):
    lib = _lib.lib()
    h, w = flow2.shape[2:]
    out = p.f32(B, 2, 4 * h, 4 * w)
    p.call("upsample4x", lib.ft_upsample4x, flow2.data_ptr(), out.data_ptr(), B, 2, h, w, float(mul), 0 if mode == "bilinear" else 1)
    return out


def _frames(p, xf, B, H, W, xpad=(0, 0)):
    """[B,6,H,W] fp32 -> split-fp16 NHWC [2B,H,W,8]: first frames then second frames (FlowNetC's two streams)."""
    lib = _lib.lib()
    fr = p.act(2 * B, H, W, 3, xpad=xpad)
    for f in range(2):
        p.call("pack_image", lib.ft_pack_image, xf.data_ptr() + f * 3 * H * W * 4, 6 * H * W, H * W, W, 1, 3, 1.0, fr.view(0, 3, f * B, B).struct(),
               xpad[0], xpad[1])
    return fr


def _warp_concat(p, xf, flow, B, H, W, div_flow):
    lib = _lib.lib()
    cat = p.act(B, H, W, 12)
    p.call("warp_concat", lib.ft_warp_concat, xf.data_ptr(), flow.data_ptr(), float(div_flow), B, H, W, cat.struct())
    return cat


class FlowNet2S(FlowNetS.FlowNetS, _FlowBase):
    def __init__(self, args=None, batchNorm=False, div_flow=20):
        FlowNetS.FlowNetS.__init__(self, args, input_channels=6, batchNorm=batchNorm)
        self._init_base(args, div_flow)
        self.eval()

    def _build(self, p, L, pair, B, H, W):
        p.rgb_max = self.rgb_max
        _, xa = _normalise(p, pair, B, H, W, False, True, xpad=L[""]["conv1"].in_xpad)
        return _up4(p, FlowNetS.build(p, L[""], xa, B), B, self.div_flow, "bilinear")


class FlowNet2C(FlowNetC.FlowNetC, _FlowBase):
    def __init__(self, args=None, batchNorm=False, div_flow=20):
        FlowNetC.FlowNetC.__init__(self, args, batchNorm=batchNorm, div_flow=20)
        self._init_base(args, 20)
        self.eval()

    def _build(self, p, L, pair, B, H, W):
        p.rgb_max = self.rgb_max
        xf, _ = _normalise(p, pair, B, H, W, True, False)
        return _up4(p, FlowNetC.build(p, L[""], _frames(p, xf, B, H, W, L[""]["conv1"].in_xpad), B), B, self.div_flow, "bilinear")


class FlowNet2SD(FlowNetSD.FlowNetSD, _FlowBase):
    def __init__(self, args=None, batchNorm=False, div_flow=20):
        FlowNetSD.FlowNetSD.__init__(self, args, batchNorm=batchNorm)
        self._init_base(args, div_flow)
        self.eval()

    def _build(self, p, L, pair, B, H, W):
        p.rgb_max = self.rgb_max
        _, xa = _normalise(p, pair, B, H, W, False, True, xpad=L[""]["conv0"].in_xpad)
        return _up4(p, FlowNetSD.build(p, L[""], xa, B), B, self.div_flow, "bilinear")


class _Stacked(_FlowBase):
    def _subnets(self):
        return {n: m for n, m in self.named_children() if n.startswith("flownet")}

    def _c_then_s1(self, p, L, pair, B, H, W, sd_branch=False):
        """FlowNetC -> warp -> FlowNetS_1 (models.py:121-135); returns x (fp32), x (act) and the 2nd full-res flow.
        sd_branch (FlowNet2): FlowNetSD only needs the normalised pair, so it is added here as a branch of the plan that runs
        next to the whole C -> S1 -> S2 chain (joined before the fusion network)."""
        p.rgb_max = self.rgb_max
        xf, xa = _normalise(p, pair, B, H, W, True, True, xpad=L["flownets_d"]["conv0"].in_xpad if "flownets_d" in L else (0, 0))
        if sd_branch:
            with p.branch() as br:
                self._sd_flow = _up4(p, FlowNetSD.build(p, L["flownets_d"], xa, B, "sd_"), B, 1.0 / self.div_flow, "nearest")
            self._sd_branch = br
        flow = _up4(p, FlowNetC.build(p, L["flownetc"], _frames(p, xf, B, H, W, L["flownetc"]["conv1"].in_xpad), B, "c_"), B, self.div_flow, "bilinear")
        p.taps["flownetc_flow"] = flow
        cat = _warp_concat(p, xf, flow, B, H, W, self.div_flow)
        flow = _up4(p, FlowNetS.build(p, L["flownets_1"], cat, B, "s1_"), B, self.div_flow, "bilinear")
        p.taps["flownets1_flow"] = flow
        return xf, xa, flow


class FlowNet2CS(_Stacked):
    def __init__(self, args=None, batchNorm=False, div_flow=20.):
        nn.Module.__init__(self)
        self.batchNorm = batchNorm
        self.channelnorm, self.resample1 = ChannelNorm(), Resample2d()
        self.flownetc = FlowNetC.FlowNetC(args, batchNorm=batchNorm)
        self.flownets_1 = FlowNetS.FlowNetS(args, batchNorm=batchNorm)
        self.upsample1, self.upsample2 = nn.Upsample(scale_factor=4, mode="bilinear"), nn.Upsample(scale_factor=4, mode="bilinear")
        sm.flow_init(self)
        self._init_base(args, div_flow)
        self.eval()

    def _build(self, p, L, pair, B, H, W):
        return self._c_then_s1(p, L, pair, B, H, W)[2]


class FlowNet2CSS(_Stacked):
    def __init__(self, args=None, batchNorm=False, div_flow=20.):
        nn.Module.__init__(self)
        self.batchNorm = batchNorm
        self.channelnorm, self.resample1, self.resample2 = ChannelNorm(), Resample2d(), Resample2d()
        self.flownetc = FlowNetC.FlowNetC(args, batchNorm=batchNorm)
        self.flownets_1 = FlowNetS.FlowNetS(args, batchNorm=batchNorm)
        self.flownets_2 = FlowNetS.FlowNetS(args, batchNorm=batchNorm)
        self.upsample1, self.upsample2 = nn.Upsample(scale_factor=4, mode="bilinear"), nn.Upsample(scale_factor=4, mode="bilinear")
        self.upsample3 = nn.Upsample(scale_factor=4, mode="nearest")
        sm.flow_init(self)
        self._init_base(args, div_flow)
        self.eval()

    def _build(self, p, L, pair, B, H, W):
        xf, xa, flow = self._c_then_s1(p, L, pair, B, H, W)
        cat = _warp_concat(p, xf, flow, B, H, W, self.div_flow)
        return _up4(p, FlowNetS.build(p, L["flownets_2"], cat, B, "s2_"), B, self.div_flow, "nearest")


class FlowNet2(_Stacked):
    def __init__(self, args=None, batchNorm=False, div_flow=20.):
        nn.Module.__init__(self)
        self.batchNorm = batchNorm
        self.channelnorm = ChannelNorm()
        self.resample1, self.resample2, self.resample3, self.resample4 = Resample2d(), Resample2d(), Resample2d(), Resample2d()
        self.flownetc = FlowNetC.FlowNetC(args, batchNorm=batchNorm)
        self.flownets_1 = FlowNetS.FlowNetS(args, batchNorm=batchNorm)
        self.flownets_2 = FlowNetS.FlowNetS(args, batchNorm=batchNorm)
        self.flownets_d = FlowNetSD.FlowNetSD(args, batchNorm=batchNorm)
        self.flownetfusion = FlowNetFusion.FlowNetFusion(args, batchNorm=batchNorm)
        self.upsample1, self.upsample2 = nn.Upsample(scale_factor=4, mode="bilinear"), nn.Upsample(scale_factor=4, mode="bilinear")
        self.upsample3, self.upsample4 = nn.Upsample(scale_factor=4, mode="nearest"), nn.Upsample(scale_factor=4, mode="nearest")
        sm.flow_init(self)
        self._init_base(args, div_flow)
        self.eval()

    def _build(self, p, L, pair, B, H, W):
        lib = _lib.lib()
        xf, xa, flow = self._c_then_s1(p, L, pair, B, H, W, sd_branch=True)
        cat = _warp_concat(p, xf, flow, B, H, W, self.div_flow)
        s2 = _up4(p, FlowNetS.build(p, L["flownets_2"], cat, B, "s2_"), B, self.div_flow, "nearest")            # models.py:146-147
        sd = self._sd_flow                                                                                       # models.py:159-160 ('/'), built as a branch
        self._sd_branch.join()
        p.taps["flownets2_flow"], p.taps["flownetsd_flow"] = s2, sd
        fcat = p.act(B, H, W, 11)
        p.call("fusion_concat", lib.ft_fusion_concat, xf.data_ptr(), sd.data_ptr(), s2.data_ptr(), B, H, W, fcat.struct())
        return FlowNetFusion.build(p, L["flownetfusion"], fcat, B, p.f32(B, 2, H, W))
