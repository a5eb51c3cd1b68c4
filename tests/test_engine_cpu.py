"""CPU (no GPU): host-side plumbing of the B200 path -- BN folding, GEMM weight ordering, the 4-phase
deconv decomposition, the tcgen05 tile image, state_dict compatibility, library symbols (config 1)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch
import torch.nn as nn

from flowtrack.pytorch_b200 import _lib, engine
from tests import emulate
from tests.golden import recipes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rand_bn(c, g):
    bn = nn.BatchNorm2d(c).eval()
    with torch.no_grad():
        bn.weight.copy_(torch.rand(c, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(c, generator=g) * 0.1)
        bn.running_mean.copy_(torch.randn(c, generator=g) * 0.2)
        bn.running_var.copy_(torch.rand(c, generator=g) + 0.5)
    return bn


@pytest.mark.parametrize("cin,cout,k,stride,hw", [(3, 64, 7, 2, (32, 24)), (64, 64, 1, 1, (8, 6)), (24, 40, 3, 2, (9, 7)),
                                                   (16, 8, 5, 2, (12, 16)), (10, 2, 3, 1, (6, 8))])
def test_conv_fold_and_ordering(cin, cout, k, stride, hw):
    g = torch.Generator().manual_seed(cin * 100 + cout)
    conv = nn.Conv2d(cin, cout, k, stride, (k - 1) // 2, bias=False)
    bn = _rand_bn(cout, g)
    layer = engine.conv_from_modules(conv, bn, _lib.ACT_RELU, keep_host=True)
    x = torch.randn(2, cin, *hw, generator=g)
    with torch.no_grad():
        ref = torch.relu(bn(conv(x)))
    out = emulate.emulate_layer(layer, emulate.nhwc(x, layer.cin_p)).permute(0, 3, 1, 2)
    assert out.shape == ref.shape
    assert (out - ref).abs().max() < 1e-5


@pytest.mark.parametrize("cin,cout,hw,bias", [(2048 // 32, 32, (8, 6), False), (10, 16, (5, 7), True), (2, 2, (3, 4), False)])
def test_deconv_four_phase_decomposition(cin, cout, hw, bias):
    g = torch.Generator().manual_seed(cin)
    dc = nn.ConvTranspose2d(cin, cout, 4, 2, 1, bias=bias)
    layer = engine.conv_from_modules(dc, None, _lib.ACT_LEAKY, keep_host=True)
    x = torch.randn(2, cin, *hw, generator=g)
    with torch.no_grad():
        ref = torch.nn.functional.leaky_relu(dc(x), 0.1)
    out = emulate.emulate_layer(layer, emulate.nhwc(x, layer.cin_p)).permute(0, 3, 1, 2)
    assert out.shape == ref.shape
    assert (out - ref).abs().max() < 1e-5


def test_tile_image_roundtrip_and_split_precision():
    conv = nn.Conv2d(72, 130, 3, 1, 1)          # FT_K_TAP64: 72 -> 128 channels per tap, K = 9*128 = 1152; cout 130 -> block 128, cout_pad 256
    layer = engine.conv_from_modules(conv, None, keep_host=True)
    assert layer.k_mode == _lib.K_TAP64 and layer.cin_k == 128
    assert layer.block_n == 128 and layer.cout_pad == 256 and layer.k_pad == 1152
    wg = layer.host["w_gemm"]
    back = emulate.untile(layer)
    rel = np.abs(back - wg).max() / np.abs(wg).max()
    assert rel < 2.0 ** -21                                # hi+lo carries >= 21 significant bits
    assert layer.w_tile_host.size == 1 * 2 * 18 * 2 * 128 * 64 * 2
    m = np.abs(wg).max(axis=(0, 1))[:130]
    assert np.all((m >= 1024) & (m < 2048))               # power-of-two column normalisation
    stem = engine.conv_from_modules(nn.Conv2d(3, 64, 7, 2, 3, bias=False), None, cin_p=8, keep_host=True)
    assert stem.k_mode == _lib.K_ROW and stem.k_pad == 7 * 64 and stem.row_pad == 8 and stem.in_xpad == (4, 4)   # one k-block per kernel row
    stem = engine.conv_from_modules(nn.Conv2d(3, 64, 7, 2, 3, bias=False), None, cin_p=8, keep_host=True, k_mode=_lib.K_PACKED)
    assert stem.k_mode == _lib.K_PACKED and stem.cin_k == 8 and stem.k_pad == 448      # K packed across taps: 49*8 = 392 -> 448
    sd0 = engine.conv_from_modules(nn.Conv2d(6, 64, 3, 1, 1), None, keep_host=True)
    assert sd0.k_mode == _lib.K_PACKED and sd0.in_xpad == (0, 0)


def test_residual_identity_blocks():
    """A layer packed for a residual call: BatchNorm's scale is folded into the weights, 1/scale is a power of two f,
    and every (phase, n_tile) of the tile image ends with ceil(block_n/64) k-blocks holding diag(f)."""
    g = torch.Generator().manual_seed(3)
    conv, bn = nn.Conv2d(64, 256, 1, bias=False), _rand_bn(256, g)
    layer = engine.conv_from_modules(conv, bn, _lib.ACT_RELU, keep_host=True, residual=True)
    assert layer.res_blocks == 2 and layer.block_n == 128
    sc = layer.host["scale"]
    f = 1.0 / sc[:256]
    assert np.all(np.log2(f) == np.round(np.log2(f))) and np.all(f <= 2.0 ** 14)
    main, resb = emulate.untile(layer, with_residual_blocks=True)
    assert np.abs(main - layer.host["w_gemm"]).max() / np.abs(main).max() < 2.0 ** -21
    for nt in range(2):
        np.testing.assert_array_equal(resb[0, nt], np.diag(f[nt * 128:(nt + 1) * 128]))
    x, r = torch.randn(2, 64, 8, 6, generator=g), torch.randn(2, 256, 8, 6, generator=g)
    with torch.no_grad():
        ref = torch.relu(bn(conv(x)) + r)
    out = emulate.emulate_layer(layer, emulate.nhwc(x, 64), res=emulate.nhwc(r, 256)).permute(0, 3, 1, 2)
    assert (out - ref).abs().max() < 1e-5


def test_pose_state_dict_layout_matches_reference(golden):
    import hashlib
    from flowtrack.pytorch_b200.pose import models
    for depth in (50, 101, 152):
        m = models.deconv("resnet%d" % depth, 17, False)
        dig = hashlib.sha256("\n".join("%s %s" % (k, tuple(v.shape)) for k, v in m.state_dict().items()).encode()).hexdigest()
        assert dig == str(golden["pose%d_keys" % depth])


def test_pose_config1_plumbing_on_cpu(golden):
    """BASELINE config 1: R50 single crop on CPU -- oracle forward vs the folded/decomposed graph the GPU path executes."""
    from flowtrack.pytorch_b200.pose import models
    orc = recipes.pose_oracle_model(50)
    m = models.deconv("resnet50", 17, False)
    m.load_state_dict(orc.state_dict(), strict=True)
    x = recipes.pose_input(50)[:1]
    with torch.no_grad():
        ref = orc(x)
    out = emulate.emulate_pose(m, x)
    assert out.shape == ref.shape == (1, 17, 32, 24)
    assert (out - ref).abs().max() < 1e-4
    assert (ref - torch.from_numpy(golden["pose50_heatmap"][:1])).abs().max() < 2e-4


def test_no_cpu_fallback():
    from flowtrack.pytorch_b200.pose import models
    m = models.deconv("resnet50", 17, False)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 3, 64, 64))
    with pytest.raises(RuntimeError, match="inference path only"):
        m.train()


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "flowtrack_b200.h")).read()
    declared = set(re.findall(r"\b(ft_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"ft_act", "ft_conv_desc"}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    l = _lib.lib()                                         # loads the .so and resolves every symbol
    assert l.ft_abi_version() == _lib.ABI_VERSION == 18
    assert l.ft_sizeof(0) == ctypes.sizeof(_lib.FtAct) == 40 and l.ft_sizeof(1) == ctypes.sizeof(_lib.FtConvDesc)


@pytest.mark.parametrize("name", recipes.FLOW_NAMES)
def test_flow_state_dict_layout_and_plan_builds_on_cpu(golden, name):
    """Reference checkpoints load strict, and the launch plan (buffer shapes, channel slices of the concat
    buffers, per-layer descriptors) is consistent -- built against CPU buffers, never executed here."""
    import hashlib
    from flowtrack.pytorch_b200.flownet.model import models
    orc = recipes.flow_oracle_model(name)
    m = getattr(models, name)(None)
    m.load_state_dict(orc.state_dict(), strict=True)
    dig = hashlib.sha256("\n".join("%s %s" % (k, tuple(v.shape)) for k, v in m.state_dict().items()).encode()).hexdigest()
    assert dig == str(golden["flow_%s_keys" % name])
    if name in ("FlowNet2CSS", "FlowNet2"):
        return                                             # same builders as CS + SD + Fusion; packing 160 M weights on CPU is slow
    p = engine.Plan("cpu")
    out = m._build(p, m._pack("cpu"), p.f32(1, 3, 2, 64, 128), 1, 64, 128)
    assert tuple(out.shape) == (1, 2, 64, 128)
    assert len(p.calls) >= 25
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 3, 2, 64, 128))


def test_fusion_plan_builds_on_cpu():
    from flowtrack.pytorch_b200.flownet.networks import FlowNetFusion, submodules
    net = FlowNetFusion.FlowNetFusion(None, batchNorm=False)
    p = engine.Plan("cpu")
    out = FlowNetFusion.build(p, submodules.pack_all(net, "cpu"), p.act(1, 64, 128, 11), 1, p.f32(1, 2, 64, 128))
    assert len(p.calls) == 17 and tuple(out.shape) == (1, 2, 64, 128)


@pytest.mark.parametrize("shape", [(1, 8, 13, 31), (2, 4, 16, 24), (1, 4, 2, 2), (1, 4, 50, 20)])
def test_correlation_banded_gemm_decomposition(shape):
    """The tile / chunk / band bookkeeping of the tcgen05 correlation kernel (csrc/ft_corr.cu), emulated on the CPU,
    reproduces the reference cost volume (every output written exactly once, zeros outside the image)."""
    from oracle import ops_oracle
    g = torch.Generator().manual_seed(5)
    a, b = torch.randn(*shape, generator=g), torch.randn(*shape, generator=g)
    ref = ops_oracle.correlation_torch(a, b, 20, 1, 20, 1, 2).double()
    out = emulate.emulate_corr_banded(a, b)
    assert not torch.isnan(out).any()
    assert (out - ref).abs().max() < 1e-6


@pytest.mark.parametrize("H", [2, 6, 13, 16, 48, 50, 96])
@pytest.mark.parametrize("parts", [1, 2, 3, 5])
def test_correlation_pair_units_cover_every_displacement_once(H, parts):
    """The CTA-pair work decomposition of the correlation kernel (corr_pair_unit): every (in1' row, displacement row) of
    every parity class is produced exactly once -- by a chunk when the in2' row exists, by the zero loop when it does not."""
    for py, (H2c, cover) in emulate.corr_pair_unit_coverage(H, parts).items():
        want = {(a, tj) for a in range(H2c) for tj in range(-10, 11)}
        assert set(cover) == want, (py, sorted(want - set(cover))[:5], sorted(set(cover) - want)[:5])
        assert all(v == 1 for v in cover.values())


def test_stem_space_to_depth_transform():
    """7x7 stride-2 pad-3 stem == 4x4 stride-1 (pad 2 before, 1 after) conv over the space-to-depth image: the weight
    transform in plain torch, and the packed FT_K_ROW layer (4 k-blocks) through the CPU emulation of ft_conv2d_fprop."""
    torch.manual_seed(0)
    conv = nn.Conv2d(3, 64, 7, 2, 3, bias=False).double()
    bn = nn.BatchNorm2d(64).double().eval()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5), bn.bias.normal_(0, 0.1), bn.running_mean.normal_(0, 0.1), bn.running_var.uniform_(0.75, 1.25)
    x = torch.randn(2, 3, 32, 24, dtype=torch.float64)
    ref = torch.relu(bn(conv(x)))
    s2d = torch.zeros(2, 16, 16, 12, dtype=torch.float64)
    for py in range(2):
        for px in range(2):
            s2d[:, (py * 2 + px) * 4:(py * 2 + px) * 4 + 3] = x[:, :, py::2, px::2]
    w4 = engine.stem_s2d_weights(conv.weight).double()
    assert (torch.nn.functional.conv2d(torch.nn.functional.pad(s2d, (2, 1, 2, 1)), w4) - conv(x)).abs().max() < 1e-12
    layer = engine.stem_s2d_from_modules(conv.float(), bn.float(), _lib.ACT_RELU, device=None, keep_host=True)
    assert layer.k_mode == _lib.K_ROW and layer.k_pad == 256 and layer.out_hw(16, 12) == (16, 12) and (layer.pad, layer.pad_hi) == (2, 1)
    assert layer.flops(2, 16, 12) == 2 * 2 * 16 * 12 * 3 * 49 * 64            # algorithmic FLOPs of the 7x7 convolution
    y = emulate.emulate_layer(layer, s2d.permute(0, 2, 3, 1).float().contiguous())
    assert (y.permute(0, 3, 1, 2)[:, :64].double() - ref).abs().max() < 1e-5


@pytest.mark.parametrize("qw,qh,n", [(48, 64, 64), (6, 8, 64), (6, 8, 5), (96, 128, 3), (12, 16, 1), (150, 20, 1), (3, 3, 2)])
@pytest.mark.parametrize("n_tiles,n_phase,pair", [(1, 1, False), (2, 1, True), (8, 4, True), (3, 4, False), (16, 1, True)])
def test_conv_tile_decode_covers_every_output_once(qw, qh, n, n_tiles, n_phase, pair):
    """tile_coord (multiply-high division, pixel boxes, CTA pairs, reverse order): every (phase, n-tile, pixel) of the output is
    produced by exactly one tile, in both sweep directions."""
    for rev in (False, True):
        visited, m_tiles, cover = emulate.conv_tile_coverage(qw, qh, n, n_tiles, n_phase, pair, rev)
        real = [v for v in visited if v[1] < m_tiles]
        assert len(set(real)) == len(real) == n_phase * m_tiles * n_tiles
        assert (cover == 1).all()


def test_fastdiv_magic_is_exact():
    rng = np.random.default_rng(0)
    for d in list(range(1, 300)) + [384, 1536, 3072, 6240, 6400, 65536, 99991]:
        f = emulate.make_fastdiv(d)
        for nval in [0, 1, d - 1, d, d + 1, 2 * d - 1, 10 ** 6, (1 << 27) - 1] + rng.integers(0, 1 << 28, 40).tolist():
            assert emulate.fdiv(int(nval), f) == int(nval) // d
