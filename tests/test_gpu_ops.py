"""GPU parity: native flow ops, layout kernels and decode vs the CPU oracle (and, for Correlation,
vs the reference's own kernel compiled into oracle/_ref when that .so travelled with the snapshot)."""
import ctypes
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from flowtrack.pytorch_b200 import _lib, engine
from oracle import decode_oracle, flow_oracle, ops_oracle
from tests import gpu_util
from tests.golden import recipes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
L = _lib.lib


def _corr(a, b, pad, k, md, s1, s2, leaky=0):
    B, C, H, W = a.shape
    oc, oh, ow = ops_oracle.correlation_shape(H, W, pad, k, md, s1, s2)
    out = torch.full((B, oc, oh, ow), float("nan"), device="cuda")
    _lib.check(L().ft_correlation_forward(a.data_ptr(), b.data_ptr(), out.data_ptr(), B, C, H, W, pad, k, md, s1, s2, leaky, gpu_util.sp()), "corr")
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("shape,params", [((2, 256, 48, 64), (20, 1, 20, 1, 2)),      # FlowNetC shape (fast path)
                                          ((1, 40, 13, 70), (20, 1, 20, 1, 2)),       # ragged: H not /8, W > 64, C not /4
                                          ((2, 24, 12, 16), (4, 1, 4, 1, 1)),         # generic path
                                          ((1, 8, 14, 18), (3, 3, 2, 2, 1))])         # 3x3 patches, stride1 2
def test_correlation_vs_oracle(shape, params):
    g = torch.Generator().manual_seed(sum(shape))
    a, b = torch.randn(*shape, generator=g), torch.randn(*shape, generator=g)
    ref = ops_oracle.correlation_torch(a, b, *params)
    out = _corr(a.cuda(), b.cuda(), *params).cpu()
    assert out.shape == ref.shape
    assert torch.isfinite(out).all()
    assert (out - ref).abs().max() < 1e-5 * max(1.0, ref.abs().max().item())
    if shape[0] * shape[1] * shape[2] * shape[3] < 300000:
        assert (out - ops_oracle.correlation_c(a, b, *params)).abs().max() < 1e-5


def _corr_ws(a, b, leaky=0, precision=_lib.PREC_SPLIT):
    """ft_correlation_forward_ws: the tcgen05 banded-GEMM kernel behind the reference's fp32 NCHW signature."""
    B, C, H, W = a.shape
    out = torch.full((B, 441, H, W), float("nan"), device="cuda")
    n = L().ft_correlation_workspace_bytes(B, C, H, W)
    ws = torch.empty(n, dtype=torch.uint8, device="cuda")
    assert ws.data_ptr() % 256 == 0
    _lib.check(L().ft_correlation_forward_ws(a.data_ptr(), b.data_ptr(), out.data_ptr(), B, C, H, W, 20, 1, 20, 1, 2, leaky, precision,
                                             ws.data_ptr(), n, gpu_util.sp()), "corr_ws")
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("shape", [(2, 256, 48, 64),      # FlowNetC at 384x512
                                   (3, 256, 16, 24),      # FlowNetC at 128x192 (tiles with out-of-image rows and columns)
                                   (1, 64, 13, 31),       # odd H and W: the parity classes differ in size
                                   (2, 128, 50, 64),      # H/2 not a multiple of the 4-row tile
                                   (1, 192, 2, 2)])       # one pixel per parity class
@pytest.mark.parametrize("variant", ["4", "2", "3"], ids=["cta_pair", "single_cta", "a_in_tmem"])
def test_correlation_tensor_core_vs_oracle(shape, variant, request):
    L().ft_set_debug(3, int(variant))                        # kernel variant knob (explicit; the library never reads the environment)
    request.addfinalizer(lambda: L().ft_set_debug(3, 0))
    assert L().ft_correlation_tc_supported(shape[1], shape[2], shape[3], 20, 1, 20, 1, 2) == 1
    g = torch.Generator().manual_seed(7 + sum(shape))
    a, b = torch.randn(*shape, generator=g), torch.randn(*shape, generator=g)
    ref = ops_oracle.correlation_torch(a, b, 20, 1, 20, 1, 2)
    out = _corr_ws(a.cuda(), b.cuda()).cpu()
    assert torch.isfinite(out).all()
    err = (out - ref).abs().max().item()
    print("tcgen05 correlation %s: max-abs %.3g (ref abs-max %.3g)" % (shape, err, ref.abs().max().item()))
    assert err < 1e-5 * max(1.0, ref.abs().max().item())
    simt = _corr(a.cuda(), b.cuda(), 20, 1, 20, 1, 2).cpu()                       # the CUDA-core kernel on the same inputs
    assert (out - simt).abs().max() < 1e-5
    fast = _corr_ws(a.cuda(), b.cuda(), precision=_lib.PREC_FAST).cpu()           # fp16 operands: ~2^-11 relative per product
    assert (fast - ref).abs().max() < 4e-3 * max(1.0, ref.abs().max().item())
    ol = _corr_ws(a.cuda(), b.cuda(), leaky=1).cpu()
    assert torch.equal(ol, torch.where(out > 0, out, 0.1 * out))


@pytest.mark.parametrize("shape", [(1, 128, 24, 128),     # W = 128: two 32-column strips per parity class, four column chunks
                                   (2, 256, 96, 128),     # FlowNetC at 768x1024 frames
                                   (1, 64, 21, 100),      # odd H, W/2 = 50: ragged last strip
                                   (1, 64, 9, 66)])       # just past the single-strip width
def test_correlation_tensor_core_wide_frames_vs_oracle(shape):
    """Maps wider than 64 pixels: the CTA-pair kernel tiles every parity class in 32-column strips, each in2' row chunk as two
    column chunks (columns b0-10 .. b0+53), every band position written by exactly one of them."""
    assert L().ft_correlation_tc_supported(shape[1], shape[2], shape[3], 20, 1, 20, 1, 2) == 1
    g = torch.Generator().manual_seed(17 + sum(shape))
    a, b = torch.randn(*shape, generator=g), torch.randn(*shape, generator=g)
    ref = ops_oracle.correlation_torch(a, b, 20, 1, 20, 1, 2)
    nan = torch.full_like(ref, float("nan")).cuda()
    out = _corr_ws(a.cuda(), b.cuda())
    assert torch.isfinite(out).all()
    err = (out.cpu() - ref).abs().max().item()
    print("tcgen05 correlation (wide) %s: max-abs %.3g (ref abs-max %.3g)" % (shape, err, ref.abs().max().item()))
    assert err < 1e-5 * max(1.0, ref.abs().max().item())
    if shape[2] % 2 == 0 and shape[3] % 2 == 0:             # the parity-planar layout FlowNetC consumes
        aa, ba = gpu_util.pack_nchw(a.cuda()), gpu_util.pack_nchw(b.cuda())
        B, C, H, W = shape
        planar = torch.full((B, 441, 4, H // 2, W // 2), float("nan"), device="cuda")
        _lib.check(L().ft_correlation_act_planar(aa.struct(), ba.struct(), planar.data_ptr(), 0, _lib.PREC_SPLIT, gpu_util.sp()), "corr_planar")
        torch.cuda.synchronize()
        want = torch.stack([out[:, :, py::2, px::2] for py in (0, 1) for px in (0, 1)], dim=2)
        assert torch.equal(planar, want)
    del nan


def test_correlation_tensor_core_full_size_properties():
    """BASELINE config 3 size [32,256,48,64]: checked through size-independent properties (the oracle is too slow)."""
    g = torch.Generator(device="cuda").manual_seed(11)
    a, b = torch.randn(32, 256, 48, 64, device="cuda", generator=g), torch.randn(32, 256, 48, 64, device="cuda", generator=g)
    o = _corr_ws(a, b)
    assert torch.isfinite(o).all()
    assert (o[:, 220] - (a * b).mean(1)).abs().max() < 1e-5                        # zero displacement = mean_c a*b
    # displacement (tj, ti) = (+3, -4) -> in2 shifted by (+6, -8) pixels, zero outside the image
    sh = torch.zeros_like(b)
    sh[:, :, :42, 8:] = b[:, :, 6:, :56]
    assert (o[:, (3 + 10) * 21 + (-4 + 10)] - (a * sh).mean(1)).abs().max() < 1e-5
    assert (_corr_ws(2 * a, b) - 2 * o).abs().max() < 1e-5                         # linear in input1
    # swapping the inputs mirrors the displacement: corr(b,a)[-d](p) = corr(a,b)[d](p - 2d); check the centre row of channels
    o2 = _corr_ws(b, a)
    assert (o2[:, 220] - o[:, 220]).abs().max() < 1e-5
    assert (o2[:, (10 - 2) * 21 + 10, 4:, :] - o[:, (10 + 2) * 21 + 10, :44, :]).abs().max() < 1e-5


@pytest.mark.parametrize("shape", [(2, 256, 48, 64), (3, 128, 16, 24), (1, 64, 6, 2)])
@pytest.mark.parametrize("variant", ["2", "4"], ids=["single_cta", "cta_pair"])
def test_correlation_planar_layout_and_pack(shape, variant, request):
    """ft_correlation_act_planar (parity-planar fp32 [B,441,4,H/2,W/2]) == the NCHW result re-indexed -- from the single-CTA kernel
    (the default) and from the CTA-pair kernel -- and ft_pack_corr_planar lands it in channels [32, 473) of a 480-channel NHWC
    merge buffer like ft_pack_image does."""
    L().ft_set_debug(3, int(variant))
    request.addfinalizer(lambda: L().ft_set_debug(3, 0))
    B, C, H, W = shape
    g = torch.Generator().manual_seed(3 + sum(shape))
    a, b = torch.randn(*shape, generator=g).cuda(), torch.randn(*shape, generator=g).cuda()
    aa, ba = gpu_util.pack_nchw(a), gpu_util.pack_nchw(b)
    nchw = torch.full((B, 441, H, W), float("nan"), device="cuda")
    planar = torch.full((B, 441, 4, H // 2, W // 2), float("nan"), device="cuda")
    _lib.check(L().ft_correlation_act(aa.struct(), ba.struct(), nchw.data_ptr(), 1, _lib.PREC_SPLIT, gpu_util.sp()), "corr_act")
    _lib.check(L().ft_correlation_act_planar(aa.struct(), ba.struct(), planar.data_ptr(), 1, _lib.PREC_SPLIT, gpu_util.sp()), "corr_planar")
    torch.cuda.synchronize()
    want = torch.stack([nchw[:, :, py::2, px::2] for py in (0, 1) for px in (0, 1)], dim=2)
    assert torch.equal(planar, want)
    ref = ops_oracle.correlation_torch(a.cpu(), b.cpu(), 20, 1, 20, 1, 2)
    ref = torch.where(ref > 0, ref, 0.1 * ref)
    assert (nchw.cpu() - ref).abs().max() < 1e-5 * max(1.0, ref.abs().max().item())
    m1, m2 = engine.Act(B, H, W, 32 + 441, "cuda", zero=True), engine.Act(B, H, W, 32 + 441, "cuda", zero=True)
    st = nchw.stride()
    _lib.check(L().ft_pack_image(nchw.data_ptr(), st[0], st[1], st[2], st[3], 441, 1.0, m1.view(32, 441).struct(), 0, 0, gpu_util.sp()), "pack_image")
    _lib.check(L().ft_pack_corr_planar(planar.data_ptr(), m2.view(32, 441).struct(), gpu_util.sp()), "pack_corr_planar")
    torch.cuda.synchronize()
    assert torch.equal(m1.hi, m2.hi) and torch.equal(m1.lo, m2.lo)


def test_correlation_leaky_and_linearity():
    g = torch.Generator().manual_seed(1)
    a, b = torch.randn(2, 256, 48, 64, generator=g).cuda(), torch.randn(2, 256, 48, 64, generator=g).cuda()
    o = _corr(a, b, 20, 1, 20, 1, 2)
    ol = _corr(a, b, 20, 1, 20, 1, 2, leaky=1)
    assert torch.equal(ol, torch.where(o > 0, o, 0.1 * o))
    assert (_corr(2 * a, b, 20, 1, 20, 1, 2) - 2 * o).abs().max() < 1e-6          # linear in input1
    centre = o[:, 220]                                                             # zero displacement = mean_c a*b
    assert (centre - (a * b).mean(1)).abs().max() < 1e-5


def test_correlation_vs_reference_kernel():
    so = os.path.join(ROOT, "oracle", "_ref", "libref_correlation.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libref_correlation.so not built (needs /root/reference at build time)")
    ref = ctypes.CDLL(so)
    g = torch.Generator().manual_seed(2)
    B, C, H, W = 2, 256, 48, 64
    a, b = torch.randn(B, C, H, W, generator=g).cuda(), torch.randn(B, C, H, W, generator=g).cuda()
    out = torch.zeros(B, 441, H, W, device="cuda")
    r1 = torch.zeros(B, H + 40, W + 40, C, device="cuda")
    r2 = torch.zeros_like(r1)
    vp, ci = ctypes.c_void_p, ctypes.c_int
    ref.Correlation_forward_cuda_kernel.argtypes = [vp] + [ci] * 8 + [vp] + [ci] * 7 + [vp] + [ci] * 5 + [vp, vp] + [ci] * 6 + [vp]
    ok = ref.Correlation_forward_cuda_kernel(out.data_ptr(), B, 441, H, W, *out.stride(), a.data_ptr(), C, H, W, *a.stride(),
                                             b.data_ptr(), C, *b.stride(), r1.data_ptr(), r2.data_ptr(), 20, 1, 20, 1, 2, 1, None)
    torch.cuda.synchronize()
    assert ok == 1
    mine = _corr(a, b, 20, 1, 20, 1, 2)
    assert (mine - out).abs().max() < 1e-5


def test_resample2d_and_channelnorm():
    g = torch.Generator().manual_seed(3)
    img, flow = torch.randn(2, 3, 96, 128, generator=g), torch.randn(2, 2, 96, 128, generator=g) * 9
    flow[0, :, :4] = 500.0                                                        # far out of range: clamp-to-edge
    out = torch.empty(2, 3, 96, 128, device="cuda")
    img_d, flow_d = img.cuda(), flow.cuda()                                      # keep device copies alive across the launch
    _lib.check(L().ft_resample2d_forward(img_d.data_ptr(), flow_d.data_ptr(), out.data_ptr(), 2, 3, 96, 128, gpu_util.sp()), "resample")
    assert (out.cpu() - ops_oracle.resample2d_c(img, flow)).abs().max() < 1e-5
    zero = torch.zeros_like(flow).cuda()
    _lib.check(L().ft_resample2d_forward(img_d.data_ptr(), zero.data_ptr(), out.data_ptr(), 2, 3, 96, 128, gpu_util.sp()), "resample")
    assert torch.equal(out.cpu(), img)                                            # identity warp is exact
    x = torch.randn(2, 3, 96, 128, generator=g)
    x_d = x.cuda()
    n = torch.empty(2, 1, 96, 128, device="cuda")
    _lib.check(L().ft_channelnorm_forward(x_d.data_ptr(), n.data_ptr(), 2, 3, 96, 128, gpu_util.sp()), "cnorm")
    assert (n.cpu() - ops_oracle.channelnorm_c(x)).abs().max() < 1e-6


@pytest.mark.parametrize("mode", [0, 1])
def test_upsample4x(mode):
    x = torch.randn(2, 2, 24, 32, generator=torch.Generator().manual_seed(4))
    out = torch.empty(2, 2, 96, 128, device="cuda")
    x_d = x.cuda()
    _lib.check(L().ft_upsample4x(x_d.data_ptr(), out.data_ptr(), 2, 2, 24, 32, 20.0, mode, gpu_util.sp()), "up")
    ref = F.interpolate(x * 20.0, scale_factor=4, mode="bilinear", align_corners=False) if mode == 0 else F.interpolate(x * 20.0, scale_factor=4, mode="nearest")
    assert (out.cpu() - ref).abs().max() < 1e-4


def test_pack_pair_and_layout_roundtrip():
    pair = recipes.flow_input(2, (64, 128))
    B, H, W = 2, 64, 128
    d = engine.Act(B, H, W, 8, "cuda")
    xf = torch.empty(B, 6, H, W, device="cuda")
    ws = torch.empty(L().ft_pack_pair_ws_floats(B), device="cuda")
    pair_d = pair.cuda()
    _lib.check(L().ft_pack_pair(pair_d.data_ptr(), B, H, W, 255.0, ws.data_ptr(), xf.data_ptr(), d.struct(), 0, 0, gpu_util.sp()), "pack_pair")
    ref = flow_oracle.normalise_pair(pair, 255.0)
    assert (xf.cpu() - ref).abs().max() < 1e-6
    back = d.to_nchw().cpu()
    assert (back[:, :6] - ref).abs().max() < 1e-6 and back[:, 6:].abs().max() == 0   # split-fp16 keeps ~22 bits


def test_maxpool():
    x = torch.randn(2, 64, 32, 24, generator=torch.Generator().manual_seed(6)).cuda()
    a = gpu_util.pack_nchw(x)
    o = engine.Act(2, 16, 12, 64, "cuda")
    _lib.check(L().ft_maxpool3x3s2(a.struct(), o.struct(), gpu_util.sp()), "maxpool")
    assert torch.equal(o.to_nchw(), F.max_pool2d(a.to_nchw(), 3, 2, 1))


def test_decode_matches_oracle_and_golden(golden):
    hm = recipes.decode_heatmaps()
    center = torch.tensor(golden["decode_center"], dtype=torch.float32).cuda()
    scale = torch.tensor(golden["decode_scale"], dtype=torch.float32).cuda()
    hm_d = hm.cuda()
    for adjust, key in ((1, "decode_coords_adjust"), (0, "decode_coords_plain")):
        kps = torch.empty(3, 17, 3, device="cuda")
        _lib.check(L().ft_decode_keypoints(hm_d.data_ptr(), 3, 17, 64, 48, center.data_ptr(), scale.data_ptr(), adjust, 0.0, None, kps.data_ptr(), gpu_util.sp()), "decode")
        kps = kps.cpu().numpy()
        np.testing.assert_allclose(kps[:, :, :2], golden[key], atol=2e-3)          # fixtures made by the real final_preds
        np.testing.assert_array_equal(kps[:, :, 2:], golden["decode_scores"])
        c, s = decode_oracle.final_preds(hm.numpy(), golden["decode_center"], golden["decode_scale"], bool(adjust))
        np.testing.assert_allclose(kps[:, :, :2], c, atol=2e-3)
    # ties: first maximum wins (torch.max semantics)
    flat = torch.zeros(1, 1, 4, 6)
    flat[0, 0, 1, 2] = flat[0, 0, 3, 5] = 1.0
    kps = torch.empty(1, 1, 3, device="cuda")
    one = torch.tensor([[3.0, 2.0]]).cuda()
    sc = torch.tensor([4.0]).cuda()                                                # s = h/scale = 1 -> identity + offset
    flat_d = flat.cuda()
    _lib.check(L().ft_decode_keypoints(flat_d.data_ptr(), 1, 1, 4, 6, one.data_ptr(), sc.data_ptr(), 0, 0.0, None, kps.data_ptr(), gpu_util.sp()), "decode")
    assert kps.cpu().tolist() == [[[2.0, 1.0, 1.0]]]


def test_box_propagation_vs_oracle():
    g = torch.Generator().manual_seed(8)
    B, H, W, N, K = 3, 96, 128, 7, 17
    flow = torch.randn(B, 2, H, W, generator=g) * 5
    kp = torch.rand(N, K, 3, generator=g) * torch.tensor([W - 1.0, H - 1.0, 1.0])
    kp[:, ::5, 2] = -0.2                                                           # some joints below threshold
    frame_of = torch.tensor([0, 0, 1, 1, 2, 2, 2], dtype=torch.int32)
    boxes = torch.empty(N, 4, device="cuda")
    kp_d, fo_d, flow_d = kp.cuda(), frame_of.cuda(), flow.cuda()
    _lib.check(L().ft_box_propagation(kp_d.data_ptr(), fo_d.data_ptr(), N, K, flow_d.data_ptr(), B, H, W, boxes.data_ptr(), gpu_util.sp()), "boxprop")
    for n in range(N):
        ref = decode_oracle.box_propagation(kp[n:n + 1].numpy(), flow[frame_of[n]].numpy())
        np.testing.assert_allclose(boxes[n].cpu().numpy(), ref[0], atol=1e-4)


def test_bad_arguments_raise():
    with pytest.raises(RuntimeError, match="correlation"):
        _lib.check(L().ft_correlation_forward(None, None, None, 1, 1, 1, 1, 20, 1, 20, 1, 2, 0, None), "corr")
    a = engine.Act(1, 4, 4, 8, "cuda")
    bad = a.struct()
    bad.ld = 12
    with pytest.raises(RuntimeError, match="multiples of 8"):
        _lib.check(L().ft_maxpool3x3s2(bad, a.struct(), None), "maxpool")


def test_resample2d_and_channelnorm_vs_reference_kernels():
    """ft_resample2d_forward / ft_channelnorm_forward against the reference's OWN __global__ kernels
    (Resample2d_kernel.cu:20-66, ChannelNorm_kernel.cu:19-51), compiled unmodified into oracle/_ref/libref_warp.so."""
    from oracle import ref_gpu
    if ref_gpu.warp_lib() is None:
        pytest.skip("oracle/_ref/libref_warp.so not built (needs /root/reference at build time)")
    g = torch.Generator().manual_seed(13)
    B, H, W = 2, 384, 512                                                         # FlowNet2's full-resolution warps
    img, flow = (torch.rand(B, 3, H, W, generator=g) * 255).cuda(), (torch.randn(B, 2, H, W, generator=g) * 12).cuda()
    flow[0, :, :3] = -700.0                                                       # clamp-to-edge on both sides
    flow[1, :, -3:] = 900.0
    want = ref_gpu.resample2d(img, flow)
    out = torch.empty_like(img)
    _lib.check(L().ft_resample2d_forward(img.data_ptr(), flow.data_ptr(), out.data_ptr(), B, 3, H, W, gpu_util.sp()), "resample")
    # the reference forms its four weights in double and accumulates in float; this kernel keeps float weights: <= a few ulp of 255
    err = (out - want).abs().max().item()
    assert err <= 255 * 4e-7, err
    for C in (2, 3):
        x = (torch.randn(B, C, H, W, generator=g) * 5).cuda()
        want = ref_gpu.channelnorm(x)
        n = torch.empty(B, 1, H, W, device="cuda")
        _lib.check(L().ft_channelnorm_forward(x.data_ptr(), n.data_ptr(), B, C, H, W, gpu_util.sp()), "cnorm")
        assert (n - want).abs().max().item() <= 1e-5 * want.abs().max().item()


def test_flow_to_image_matches_reference_flowlib():
    """ft_flow_to_image vs fixtures made by the reference's own flowlib.flow_to_image (tests/golden/reference_flowvis.npz) and
    vs the numpy oracle on a full-size flow.  Double-precision atan2 on the GPU may differ from libm by an ulp, which can move a
    colour byte across a floor() boundary: <= 1 LSB, on < 0.1 % of the bytes."""
    import numpy as np
    from flowtrack.pytorch_b200.flownet.utils import flow_utils
    from oracle import flowvis_oracle
    from tests.golden import recipes
    fix = np.load(os.path.join(ROOT, "tests", "golden", "reference_flowvis.npz"))
    for i, flow in enumerate(recipes.flowvis_flows()):
        got = flow_utils.flow_to_image(flow)
        want = fix["flowvis_img%d" % i]
        diff = np.abs(got.astype(np.int32) - want.astype(np.int32))
        assert got.shape == want.shape and got.dtype == np.uint8
        assert diff.max() <= 1 and (diff > 0).mean() < 1e-3, (i, diff.max(), (diff > 0).mean())
    g = torch.Generator().manual_seed(4)
    f = torch.randn(3, 2, 384, 512, generator=g) * torch.tensor([1.0, 20.0, 0.01]).view(3, 1, 1, 1)
    got = flow_utils.flow_to_image_cuda(f.cuda()).cpu().numpy()
    for b in range(3):
        want = flowvis_oracle.flow_to_image(f[b].permute(1, 2, 0).numpy())
        diff = np.abs(got[b].astype(np.int32) - want.astype(np.int32))
        assert diff.max() <= 1 and (diff > 0).mean() < 1e-3
