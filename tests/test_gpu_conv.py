"""GPU parity of ft_conv2d_fprop: SIMT cross-check kernel and the tcgen05 kernel (split and fast
precision) vs torch fp32 convolution of the same op.  Tolerances: split mode is fp32-grade
(<= 2e-5 relative to the output scale), fast mode is fp16-operand (<= 2e-3 relative)."""
import ctypes

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from flowtrack.pytorch_b200 import _lib, engine
from tests import gpu_util

pytestmark = pytest.mark.gpu

CASES = [
    # name, module ctor, input shape, act
    ("1x1_64_64", lambda: nn.Conv2d(64, 64, 1, bias=False), (2, 64, 16, 12), _lib.ACT_RELU),
    ("1x1_256_64_tail", lambda: nn.Conv2d(256, 64, 1, bias=False), (3, 256, 7, 9), _lib.ACT_RELU),        # M = 189: tail tile
    ("3x3_64", lambda: nn.Conv2d(64, 64, 3, 1, 1, bias=False), (2, 64, 16, 12), _lib.ACT_RELU),
    ("3x3_s2_128", lambda: nn.Conv2d(128, 128, 3, 2, 1, bias=False), (2, 128, 16, 12), _lib.ACT_RELU),
    ("1x1_s2_down", lambda: nn.Conv2d(256, 512, 1, 2, bias=False), (2, 256, 16, 12), _lib.ACT_NONE),
    ("stem7x7", lambda: nn.Conv2d(3, 64, 7, 2, 3, bias=False), (2, 3, 64, 48), _lib.ACT_RELU),
    ("stem7x7_s1_cin6", lambda: nn.Conv2d(6, 32, 7, 1, 3), (3, 6, 20, 28), _lib.ACT_LEAKY),                  # FT_K_ROW, stride 1, batch > 1
    ("stem7x7_flow", lambda: nn.Conv2d(6, 64, 7, 2, 3), (2, 6, 64, 128), _lib.ACT_LEAKY),                  # FlowNetS conv1
    ("flow5x5_s2", lambda: nn.Conv2d(64, 128, 5, 2, 2), (1, 64, 24, 32), _lib.ACT_LEAKY),
    ("cin1032_pred", lambda: nn.Conv2d(1026, 2, 3, 1, 1), (1, 1026, 6, 8), _lib.ACT_NONE),                 # odd concat width, Cout 2
    ("head17", lambda: nn.Conv2d(256, 17, 1), (2, 256, 16, 12), _lib.ACT_NONE),
    ("wide_2048", lambda: nn.Conv2d(512, 2048, 1, bias=False), (4, 512, 8, 6), _lib.ACT_NONE),
    ("deconv_2048", lambda: nn.ConvTranspose2d(2048, 256, 4, 2, 1, bias=False), (2, 2048, 8, 6), _lib.ACT_RELU),
    ("deconv_flow", lambda: nn.ConvTranspose2d(386, 64, 4, 2, 1), (1, 386, 12, 16), _lib.ACT_LEAKY),
    ("deconv_2to2", lambda: nn.ConvTranspose2d(2, 2, 4, 2, 1, bias=False), (2, 2, 6, 8), _lib.ACT_NONE),
    # pixel boxes that span several images (8x6 and 6x8 maps), ragged in every box dimension
    ("l4_3x3_batch", lambda: nn.Conv2d(128, 64, 3, 1, 1, bias=False), (11, 128, 8, 6), _lib.ACT_RELU),
    ("l4_3x3_s2_batch", lambda: nn.Conv2d(64, 128, 3, 2, 1, bias=False), (5, 64, 16, 12), _lib.ACT_RELU),
    ("deconv_batch", lambda: nn.ConvTranspose2d(64, 32, 4, 2, 1), (9, 64, 6, 8), _lib.ACT_LEAKY),
    ("stem3x3_cin16", lambda: nn.Conv2d(11, 64, 3, 1, 1), (1, 11, 24, 32), _lib.ACT_LEAKY),                # FlowNetFusion conv0 (packed-K gather variant)
    ("wide_map_5x5", lambda: nn.Conv2d(64, 32, 5, 1, 2), (1, 64, 20, 150), _lib.ACT_NONE),                 # box wider than tall, ragged right edge
]


def _ref(mod, x, act):
    with torch.no_grad():
        y = mod(x)
    if act == _lib.ACT_RELU:
        y = torch.relu(y)
    elif act == _lib.ACT_LEAKY:
        y = F.leaky_relu(y, 0.1)
    return y


@pytest.mark.parametrize("name,ctor,shape,act", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("impl,prec,tol", [(_lib.IMPL_SIMT, _lib.PREC_SPLIT, 2e-5), (_lib.IMPL_TCGEN05, _lib.PREC_SPLIT, 2e-5),
                                           (_lib.IMPL_TCGEN05, _lib.PREC_FAST, 3e-3)], ids=["simt", "tc_split", "tc_fast"])
def test_conv_layer(name, ctor, shape, act, impl, prec, tol):
    torch.manual_seed(hash(name) % 1000)
    mod = ctor().double()
    x = torch.randn(*shape, dtype=torch.float64)
    ref = _ref(mod, x, act).float()
    layer = engine.conv_from_modules(mod.float(), None, act, device="cuda")
    xa = gpu_util.pack_nchw(x.float().cuda(), layer.cin_p, xpad=layer.in_xpad)
    out, f32 = gpu_util.run_layer(layer, xa, precision=prec, impl=impl, want_f32=True)
    got = out.to_nchw(layer.cout).cpu()
    scale = max(1.0, ref.abs().max().item())
    if impl == _lib.IMPL_TCGEN05 and prec == _lib.PREC_SPLIT and layer.k_pad > 4096:
        tol = 6e-5
    assert got.shape == ref.shape
    assert (got - ref).abs().max() <= tol * scale, (got - ref).abs().max().item()
    assert (f32.cpu() - ref).abs().max() <= tol * scale                            # fp32 NCHW output path
    if out.c > layer.cout:
        assert out.to_nchw()[:, layer.cout:].abs().max() == 0                      # padded channels are written as zeros


@pytest.mark.parametrize("prec,tol", [(_lib.PREC_SPLIT, 2e-5), (_lib.PREC_FAST, 3e-3)])
def test_bottleneck_tail_with_bn_and_residual(prec, tol):
    """conv3 of a bottleneck: 1x1 + folded BN + residual add + ReLU in one launch, written into a channel slice."""
    g = torch.Generator().manual_seed(5)
    conv = nn.Conv2d(64, 256, 1, bias=False)
    bn = nn.BatchNorm2d(256).eval()
    with torch.no_grad():
        bn.weight.copy_(torch.rand(256, generator=g) * 0.2 + 0.1)
        bn.bias.copy_(torch.randn(256, generator=g) * 0.1)
        bn.running_mean.copy_(torch.randn(256, generator=g) * 0.1)
        bn.running_var.copy_(torch.rand(256, generator=g) + 0.5)
    x, r = torch.randn(3, 64, 16, 12, generator=g), torch.randn(3, 256, 16, 12, generator=g)
    with torch.no_grad():
        ref = torch.relu(bn(conv(x)) + r)
    layer = engine.conv_from_modules(conv, bn, _lib.ACT_RELU, device="cuda", residual=True)
    xa, ra = gpu_util.pack_nchw(x.cuda()), gpu_util.pack_nchw(r.cuda())
    big = engine.Act(3, 16, 12, 256 + 64, "cuda", zero=True)                        # write into channels [64, 320) of a wider buffer
    out = gpu_util.run_layer(layer, xa, res=ra, precision=prec, out=big.view(64, 256))
    got = out.to_nchw().cpu()
    assert (got - ref).abs().max() <= tol * max(1.0, ref.abs().max().item())
    assert big.to_nchw()[:, :64].abs().max() == 0                                   # neighbours of the slice untouched


def test_tc_matches_simt_bitwise_scale_on_large_tile_count():
    """Persistent scheduling: many tiles per CTA, both TMEM accumulator stages and all smem stages wrap."""
    torch.manual_seed(11)
    conv = nn.Conv2d(128, 256, 3, 1, 1, bias=True)
    x = torch.randn(8, 128, 64, 48)                                                 # M = 24576 -> 192 m-tiles x 2 n-tiles
    layer = engine.conv_from_modules(conv, None, _lib.ACT_LEAKY, device="cuda")
    xa = gpu_util.pack_nchw(x.cuda())
    a = gpu_util.run_layer(layer, xa, impl=_lib.IMPL_TCGEN05).to_nchw()
    b = gpu_util.run_layer(layer, xa, impl=_lib.IMPL_SIMT).to_nchw()
    torch.backends.cudnn.allow_tf32 = False
    with torch.no_grad():
        ref = F.leaky_relu(conv.cuda()(x.cuda()), 0.1)
    assert (a - b).abs().max() < 2e-5 * ref.abs().max()
    assert (a - ref).abs().max() < 2e-5 * ref.abs().max()


@pytest.mark.parametrize("cin,hw,B", [(1026, (6, 8), 2), (194, (24, 32), 3), (16, (32, 64), 1)])
def test_flow_head_1x1_plus_shifted_sum(cin, hw, B):
    """predict_flow (3x3, two outputs) run as an 18-output 1x1 GEMM + ft_flow_head equals the 3x3 convolution."""
    torch.manual_seed(cin)
    conv = nn.Conv2d(cin, 2, 3, 1, 1)
    x = torch.randn(B, cin, *hw)
    with torch.no_grad():
        ref = conv.double()(x.double()).float()
    head = engine.FlowHead(conv.float(), device="cuda")
    p = engine.Plan(torch.device("cuda"))
    xa = gpu_util.pack_nchw(x.cuda(), head.cin_p)
    flow = p.flow_head(head, xa)
    p.enqueue()
    torch.cuda.synchronize()
    assert (flow.cpu() - ref).abs().max() <= 2e-5 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("bias", [False, True])
def test_flow_upsample_deconv(bias):
    torch.manual_seed(3)
    dc = nn.ConvTranspose2d(2, 2, 4, 2, 1, bias=bias)
    f = torch.randn(3, 2, 6, 8) * 5
    with torch.no_grad():
        ref = dc(f)
    up = engine.FlowUp(dc, device="cuda")
    p = engine.Plan(torch.device("cuda"))
    big = engine.Act(3, 12, 16, 64 + 2, "cuda", zero=True)
    p.flow_up(up, f.cuda(), big.view(64, 2))
    p.enqueue()
    torch.cuda.synchronize()
    got = big.view(64, 2).to_nchw(2).cpu()
    assert (got - ref).abs().max() <= 1e-5 * max(1.0, ref.abs().max().item())
    assert big.to_nchw()[:, :64].abs().max() == 0 and big.view(64, 2).to_nchw()[:, 2:].abs().max() == 0


def test_stem_packed_gather_variant_still_matches():
    """The cp.async gather variant (FT_K_PACKED) stays available for small-Cin layers that are not 7x7 stems; force it
    on the stem shape and compare with the FT_K_ROW result."""
    torch.manual_seed(2)
    conv = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
    x = torch.randn(2, 3, 64, 48)
    with torch.no_grad():
        ref = torch.relu(conv.double()(x.double())).float()
    outs = []
    for km in (None, _lib.K_PACKED):
        layer = engine.conv_from_modules(conv.float(), None, _lib.ACT_RELU, cin_p=8, device="cuda", k_mode=km)
        assert layer.k_mode == (_lib.K_ROW if km is None else _lib.K_PACKED)
        xa = gpu_util.pack_nchw(x.cuda(), 8, xpad=layer.in_xpad)
        outs.append(gpu_util.run_layer(layer, xa).to_nchw(64).cpu())
        assert (outs[-1] - ref).abs().max() <= 2e-5 * ref.abs().max()


PAIR_CASES = [
    # name, module ctor, input shape, act, residual
    ("1x1_256_128", lambda: nn.Conv2d(256, 128, 1, bias=False), (3, 256, 16, 12), _lib.ACT_RELU, False),       # 5 m-tiles: odd tail pair
    ("3x3_128_256", lambda: nn.Conv2d(128, 256, 3, 1, 1), (2, 128, 32, 24), _lib.ACT_LEAKY, False),
    ("3x3_s2_128_256", lambda: nn.Conv2d(128, 256, 3, 2, 1), (3, 128, 16, 12), _lib.ACT_RELU, False),
    ("l4_3x3_batch", lambda: nn.Conv2d(256, 512, 3, 1, 1, bias=False), (11, 256, 8, 6), _lib.ACT_RELU, False),
    ("deconv_256", lambda: nn.ConvTranspose2d(512, 256, 4, 2, 1, bias=False), (4, 512, 8, 6), _lib.ACT_RELU, False),
    ("expand_res", lambda: nn.Conv2d(64, 256, 1, bias=False), (5, 64, 16, 12), _lib.ACT_RELU, True),
    ("cin776_deconv", lambda: nn.ConvTranspose2d(770, 128, 4, 2, 1), (1, 770, 12, 16), _lib.ACT_LEAKY, False),
    ("cout64", lambda: nn.Conv2d(64, 64, 3, 1, 1, bias=False), (2, 64, 32, 24), _lib.ACT_RELU, False),
]


@pytest.mark.parametrize("name,ctor,shape,act,residual", PAIR_CASES, ids=[c[0] for c in PAIR_CASES])
@pytest.mark.parametrize("bn", [64, 128, 256])
@pytest.mark.parametrize("prec,tol", [(_lib.PREC_SPLIT, 2e-5), (_lib.PREC_FAST, 3e-3)], ids=["split", "fast"])
def test_conv_cta_pair_tiles(name, ctor, shape, act, residual, bn, prec, tol):
    """tcgen05 cta_group::2 tiles (two CTAs per 256-pixel x bn tile) against torch, every legal tile width."""
    torch.manual_seed(len(name) * 7 + bn)
    mod = ctor().double()
    x = torch.randn(*shape, dtype=torch.float64)
    with torch.no_grad():
        ref = mod(x)
    layer = engine.conv_from_modules(mod.float(), None, act, device="cuda", residual=residual)
    if layer.cout_pad % bn:
        pytest.skip("tile width does not divide the padded channel count")
    xa = gpu_util.pack_nchw(x.float().cuda(), layer.cin_p)
    ra = None
    with torch.no_grad():
        if residual:
            r = torch.randn(shape[0], layer.cout, shape[2], shape[3], dtype=torch.float64)
            ref = ref + r
            ra = gpu_util.pack_nchw(r.float().cuda())
        ref = (torch.relu(ref) if act == _lib.ACT_RELU else (F.leaky_relu(ref, 0.1) if act == _lib.ACT_LEAKY else ref)).float()
    out, f32 = gpu_util.run_layer(layer, xa, res=ra, precision=prec, want_f32=True, tiling=(bn, 1))
    got = out.to_nchw(layer.cout).cpu()
    scale = max(1.0, ref.abs().max().item())
    assert (got - ref).abs().max() <= tol * scale, (got - ref).abs().max().item()
    assert (f32.cpu() - ref).abs().max() <= tol * scale
    single = gpu_util.run_layer(layer, xa, res=ra, precision=prec, tiling=(layer.block_n, 0)).to_nchw(layer.cout).cpu()
    assert (got - single).abs().max() <= (2e-6 if prec == _lib.PREC_SPLIT else 1e-5) * scale     # same arithmetic, different tiling


@pytest.mark.parametrize("name,ctor,shape,act", [c for c in CASES if c[0] in ("3x3_s2_128", "stem7x7", "deconv_batch", "l4_3x3_batch", "stem3x3_cin16", "wide_2048")],
                         ids=["3x3_s2_128", "stem7x7", "wide_2048", "l4_3x3_batch", "deconv_batch", "stem3x3_cin16"])
def test_reverse_tile_order_is_bit_identical(name, ctor, shape, act):
    """ft_conv_desc.tile_order only changes the order in which the persistent CTAs visit the tiles (L2 reuse)."""
    torch.manual_seed(1)
    mod = ctor()
    x = torch.randn(*shape)
    layer = engine.conv_from_modules(mod, None, act, device="cuda")
    xa = gpu_util.pack_nchw(x.cuda(), layer.cin_p, xpad=layer.in_xpad)
    a, fa = gpu_util.run_layer(layer, xa, want_f32=True, tile_order=0)
    b, fb = gpu_util.run_layer(layer, xa, want_f32=True, tile_order=1)
    assert torch.equal(fa, fb) and torch.equal(a.hi, b.hi) and torch.equal(a.lo, b.lo)


@pytest.mark.parametrize("shape", [(2, 3, 64, 48), (5, 3, 96, 160)])
def test_stem_space_to_depth_path(shape):
    """ft_pack_image_s2d + the 4x4 stride-1 FT_K_ROW stem (pad 2 / pad_hi 1) == Conv2d(3, 64, 7, 2, 3) + BN + ReLU."""
    torch.manual_seed(3)
    conv = nn.Conv2d(3, 64, 7, 2, 3, bias=False).double()
    bn = nn.BatchNorm2d(64).double().eval()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5), bn.bias.normal_(0, 0.1), bn.running_mean.normal_(0, 0.1), bn.running_var.uniform_(0.75, 1.25)
    x = torch.randn(*shape, dtype=torch.float64)
    ref = torch.relu(bn(conv(x))).float()
    layer = engine.stem_s2d_from_modules(conv.float(), bn.float(), _lib.ACT_RELU, device="cuda")
    B, _, H, W = shape
    xd = x.float().cuda().contiguous()
    xa = engine.Act(B, H // 2, W // 2, 16, "cuda", xpad=layer.in_xpad)
    _lib.check(_lib.lib().ft_pack_image_s2d(xd.data_ptr(), B, H, W, xa.struct(), xa.xl, xa.xr, gpu_util.sp()), "pack_image_s2d")
    torch.cuda.synchronize()
    packed = xa.to_nchw()                                            # [B,16,H/2,W/2]
    for py in range(2):
        for px in range(2):
            c0 = (py * 2 + px) * 4
            assert (packed[:, c0:c0 + 3].cpu() - x[:, :, py::2, px::2].float()).abs().max() < 1e-6
            assert packed[:, c0 + 3].abs().max() == 0
    for impl, tol in ((_lib.IMPL_SIMT, 2e-5), (_lib.IMPL_TCGEN05, 2e-5)):
        out, f32 = gpu_util.run_layer(layer, xa, impl=impl, want_f32=True)
        scale = max(1.0, ref.abs().max().item())
        assert (f32.cpu() - ref).abs().max() < tol * scale
        assert (out.to_nchw(64).cpu() - ref).abs().max() < tol * scale


def test_autotune_picks_tilings_and_keeps_results_bit_identical(monkeypatch):
    """engine.Plan.autotune: per-layer (block_n, cta_pair) chosen by timing; the forward is bit-identical to the untuned plan."""
    from flowtrack.pytorch_b200.pose import models as pose_models
    from flowtrack.pytorch_b200.tracking import synthetic
    torch.manual_seed(0)
    m = synthetic.condition_pose_weights(pose_models.deconv("resnet50", 17, False), 0).cuda()
    x = torch.randn(4, 3, 128, 96, generator=torch.Generator().manual_seed(2)).cuda()
    monkeypatch.setenv("FT_AUTOTUNE", "0")
    a = m(x).clone()
    monkeypatch.setenv("FT_AUTOTUNE", "1")
    m.refresh()
    b = m(x).clone()
    plan = next(iter(m._plans.values()))[0]
    assert len(plan.tuned) > 40 and all(bn in (64, 128, 256) and pair in (0, 1, 2) and us > 0 and nbuf in (1, 2) and 0 <= sk <= 8 and occ in (1, 2) for bn, pair, us, nbuf, sk, occ in plan.tuned.values())
    assert (a - b).abs().max() <= 2e-6 * a.abs().max()             # bit-identical tile shapes; split-K layers differ by fp32 rounding only


def test_autotune_cache_round_trip(monkeypatch, tmp_path):
    """FT_AUTOTUNE_CACHE: the first plan of a shape measures and records its choices (split-K included), a second plan of the same
    shape loads them without measuring and computes the same bits."""
    import json
    from flowtrack.pytorch_b200.pose import models as pose_models
    from flowtrack.pytorch_b200.tracking import synthetic
    cache = tmp_path / "tuned.json"
    monkeypatch.setenv("FT_AUTOTUNE", "1")
    monkeypatch.setenv("FT_AUTOTUNE_CACHE", str(cache))
    torch.manual_seed(0)
    m = synthetic.condition_pose_weights(pose_models.deconv("resnet50", 17, False), 0).cuda()
    x = torch.randn(2, 3, 128, 96, generator=torch.Generator().manual_seed(5)).cuda()
    a = m(x).clone()
    plan_a = next(iter(m._plans.values()))[0]
    rec = json.loads(cache.read_text())
    assert list(rec) == [plan_a.signature()] and len(rec[plan_a.signature()]) == len(plan_a.tuned)
    assert any(v[4] > 1 for v in plan_a.tuned.values())            # a 2-crop batch has fewer tiles than SMs: some layers split K
    monkeypatch.setattr(torch.cuda.CUDAGraph, "replay", None)      # a second measurement would have to replay candidate graphs
    m.use_cuda_graph = False
    m.refresh()
    b = m(x).clone()
    plan_b = next(iter(m._plans.values()))[0]
    assert plan_b is not plan_a and plan_b.tuned == plan_a.tuned
    assert torch.equal(a, b)


def test_pdl_off_gives_identical_results():
    """Programmatic dependent launch only changes when a kernel's prologue runs: FT_DBG_NO_PDL (ft_set_debug(4, 1)) is bit-identical."""
    from flowtrack.pytorch_b200.pose import models as pose_models
    from flowtrack.pytorch_b200.tracking import synthetic
    torch.manual_seed(0)
    m = synthetic.condition_pose_weights(pose_models.deconv("resnet50", 17, False), 0).cuda()
    x = torch.randn(3, 3, 128, 96, generator=torch.Generator().manual_seed(3)).cuda()
    a = m(x).clone()
    try:
        _lib.lib().ft_set_debug(4, 1)
        m.refresh()
        b = m(x).clone()
    finally:
        _lib.lib().ft_set_debug(4, 0)
    assert torch.equal(a, b)


@pytest.mark.parametrize("bn,pair", [(64, 0), (64, 1), (128, 0), (128, 1)])
def test_ping_pong_epilogue_staging_is_bit_identical(bn, pair):
    """ft_conv_desc.epilogue_buffers = 2 (two staging slabs per epilogue warp group) changes scheduling only: same bits as 1,
    on a many-tile 1x1 + residual layer (every staging buffer wraps several times per CTA) and on a 3x3 layer."""
    g = torch.Generator().manual_seed(21)
    conv = nn.Conv2d(64, 256, 1, bias=False)
    bnm = nn.BatchNorm2d(256).eval()
    with torch.no_grad():
        bnm.weight.copy_(torch.rand(256, generator=g) * 0.2 + 0.1)
        bnm.running_var.copy_(torch.rand(256, generator=g) + 0.5)
    x, r = torch.randn(40, 64, 32, 24, generator=g), torch.randn(40, 256, 32, 24, generator=g)      # 240 m-tiles x 2-4 n-tiles
    layer = engine.conv_from_modules(conv, bnm, _lib.ACT_RELU, device="cuda", residual=True)
    xa, ra = gpu_util.pack_nchw(x.cuda()), gpu_util.pack_nchw(r.cuda())
    a = gpu_util.run_layer(layer, xa, res=ra, tiling=(bn, pair, 1))
    b = gpu_util.run_layer(layer, xa, res=ra, tiling=(bn, pair, 2))
    assert torch.equal(a.hi, b.hi) and torch.equal(a.lo, b.lo)
    with torch.no_grad():
        ref = torch.relu(bnm(conv(x)) + r)
    assert (b.to_nchw().cpu() - ref).abs().max() <= 2e-5 * ref.abs().max()
    conv3 = nn.Conv2d(64, 128, 3, 1, 1)
    layer3 = engine.conv_from_modules(conv3, None, _lib.ACT_LEAKY, device="cuda")
    a = gpu_util.run_layer(layer3, xa, tiling=(bn, pair, 1))
    b = gpu_util.run_layer(layer3, xa, tiling=(bn, pair, 2))
    assert torch.equal(a.hi, b.hi) and torch.equal(a.lo, b.lo)


@pytest.mark.parametrize("name,ctor,shape,act,residual", [c for c in PAIR_CASES if c[0] in ("3x3_128_256", "l4_3x3_batch", "deconv_256", "cin776_deconv", "3x3_s2_128_256")],
                         ids=["3x3_128_256", "3x3_s2_128_256", "l4_3x3_batch", "deconv_256", "cin776_deconv"])
@pytest.mark.parametrize("bn,pair,sk", [(128, 1, 2), (128, 1, 3), (64, 0, 4), (128, 0, 2), (64, 1, 8)])
def test_split_k_matches_unsplit_and_is_deterministic(name, ctor, shape, act, residual, bn, pair, sk):
    """ft_conv_desc.splitk: parts of a tile's K range on different CTAs, partial sums through the workspace, last arriver
    finalizes.  Equal to the unsplit layer up to fp32 rounding, bit-identical from run to run, counters left at zero."""
    torch.manual_seed(len(name) + sk)
    mod = ctor()
    x = torch.randn(*shape)
    layer = engine.conv_from_modules(mod, None, act, device="cuda")
    if layer.cout_pad % bn or sk * 1 > layer.k_pad // 64:
        pytest.skip("tile width / split count does not fit this layer")
    xa = gpu_util.pack_nchw(x.cuda(), layer.cin_p)
    ref = gpu_util.run_layer(layer, xa, tiling=(bn, pair)).to_nchw(layer.cout)
    _, cta_tiles = layer.tile_counts(xa.n, xa.h, xa.w, bn, pair)
    ws = (torch.empty(cta_tiles * sk * bn * 128 * 4, dtype=torch.uint8, device="cuda"), torch.zeros(cta_tiles, dtype=torch.int32, device="cuda"))
    outs = []
    for _ in range(3):
        ho, wo = layer.out_hw(xa.h, xa.w)
        out = engine.Act(xa.n, ho, wo, layer.cout, "cuda", zero=True)
        d = layer.desc(xa, out, None, None, tiling=(bn, pair, 1, sk), splitk_ws=ws)
        _lib.check(_lib.lib().ft_conv2d_fprop(ctypes.byref(d), gpu_util.sp()), "conv2d")
        torch.cuda.synchronize()
        outs.append(out)
        assert int(ws[1].abs().sum()) == 0
    got = outs[0].to_nchw(layer.cout)
    assert (got - ref).abs().max() <= 2e-6 * max(1.0, ref.abs().max().item())
    assert torch.equal(outs[0].hi, outs[1].hi) and torch.equal(outs[0].lo, outs[2].lo)


@pytest.mark.parametrize("name,ctor,shape,act,residual", PAIR_CASES, ids=[c[0] for c in PAIR_CASES])
def test_two_ctas_per_sm_variant_is_bit_identical(name, ctor, shape, act, residual):
    """ft_conv_desc.ctas_per_sm = 2 (64-wide pair tiles, two CTA pairs per TPC, 2 stages, 16-channel slabs): same bits as the
    one-CTA-per-SM kernel on every pair case, and on a many-tile stem-like layer where both resident CTAs wrap their rings."""
    torch.manual_seed(len(name) * 3)
    mod = ctor()
    x = torch.randn(*shape)
    layer = engine.conv_from_modules(mod, None, act, device="cuda", residual=residual)
    xa = gpu_util.pack_nchw(x.cuda(), layer.cin_p)
    ra = gpu_util.pack_nchw(torch.randn(shape[0], layer.cout, *layer.out_hw(shape[2], shape[3])).cuda()) if residual else None
    a = gpu_util.run_layer(layer, xa, res=ra, tiling=(64, 1, 1, 0, 1))
    b = gpu_util.run_layer(layer, xa, res=ra, tiling=(64, 1, 1, 0, 2))
    assert torch.equal(a.hi, b.hi) and torch.equal(a.lo, b.lo)


def test_two_ctas_per_sm_many_tiles_and_stem():
    torch.manual_seed(4)
    conv = nn.Conv2d(64, 64, 3, 1, 1, bias=False)
    x = torch.randn(24, 64, 64, 48)                                                  # 576 m-tiles: ~2 tiles per resident CTA
    layer = engine.conv_from_modules(conv, None, _lib.ACT_RELU, device="cuda")
    xa = gpu_util.pack_nchw(x.cuda())
    a = gpu_util.run_layer(layer, xa, tiling=(64, 1, 1, 0, 1))
    b = gpu_util.run_layer(layer, xa, tiling=(64, 1, 1, 0, 2))
    assert torch.equal(a.hi, b.hi) and torch.equal(a.lo, b.lo)
    with torch.no_grad():
        ref = torch.relu(conv(x))
    assert (b.to_nchw().cpu() - ref).abs().max() <= 2e-5 * ref.abs().max()
    stem = nn.Conv2d(6, 64, 7, 2, 3)                                                 # FlowNetS conv1 (FT_K_ROW)
    xs = torch.randn(2, 6, 128, 192)
    ls = engine.conv_from_modules(stem, None, _lib.ACT_LEAKY, device="cuda")
    xsa = gpu_util.pack_nchw(xs.cuda(), ls.cin_p, xpad=ls.in_xpad)
    a = gpu_util.run_layer(ls, xsa, tiling=(64, 1, 1, 0, 1))
    b = gpu_util.run_layer(ls, xsa, tiling=(64, 1, 1, 0, 2))
    assert torch.equal(a.hi, b.hi) and torch.equal(a.lo, b.lo)


@pytest.mark.parametrize("name,ctor,shape,act,residual", PAIR_CASES, ids=[c[0] for c in PAIR_CASES])
@pytest.mark.parametrize("bn", [64, 128])
def test_quad_clusters_with_multicast_a_tile_are_bit_identical(name, ctor, shape, act, residual, bn):
    """ft_conv_desc.cta_pair = 2: clusters of four (two CTA pairs on the same pixels, neighbouring n-tiles), the A tile fetched once
    per cluster as two multicast halves.  Same bits as plain CTA pairs on every pair case whose n-tile count is even."""
    torch.manual_seed(len(name) * 5 + bn)
    mod = ctor()
    x = torch.randn(*shape)
    layer = engine.conv_from_modules(mod, None, act, device="cuda", residual=residual)
    if layer.cout_pad % bn or (layer.cout_pad // bn) % 2:
        pytest.skip("needs an even number of n-tiles")
    xa = gpu_util.pack_nchw(x.cuda(), layer.cin_p)
    ra = gpu_util.pack_nchw(torch.randn(shape[0], layer.cout, *layer.out_hw(shape[2], shape[3])).cuda()) if residual else None
    a, fa = gpu_util.run_layer(layer, xa, res=ra, tiling=(bn, 1), want_f32=True)
    b, fb = gpu_util.run_layer(layer, xa, res=ra, tiling=(bn, 2), want_f32=True)
    assert torch.equal(a.hi, b.hi) and torch.equal(a.lo, b.lo) and torch.equal(fa, fb)


def test_quad_clusters_many_tiles():
    """Every cluster walks several work items (all stages and both TMEM buffers wrap), 8x6 maps spanning images, stride 2."""
    torch.manual_seed(9)
    for conv, shape in ((nn.Conv2d(128, 256, 3, 1, 1), (40, 128, 32, 24)), (nn.Conv2d(256, 512, 1, bias=False), (70, 256, 8, 6)),
                        (nn.Conv2d(128, 256, 3, 2, 1), (30, 128, 32, 24)), (nn.ConvTranspose2d(256, 256, 4, 2, 1, bias=False), (20, 256, 16, 12))):
        layer = engine.conv_from_modules(conv, None, _lib.ACT_RELU, device="cuda")
        xa = gpu_util.pack_nchw(torch.randn(*shape).cuda())
        a = gpu_util.run_layer(layer, xa, tiling=(128, 1))
        b = gpu_util.run_layer(layer, xa, tiling=(128, 2))
        assert torch.equal(a.hi, b.hi) and torch.equal(a.lo, b.lo)
