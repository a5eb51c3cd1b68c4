"""GPU parity: person crops (ft_crop_affine) and box NMS (ft_nms) vs the CPU oracle / the reference-generated crops."""
import os

import numpy as np
import pytest
import torch

from flowtrack.pytorch_b200 import _lib
from flowtrack.pytorch_b200.tracking import net_utils
from tests import gpu_util
from oracle import track_oracle
from tests.golden import recipes

pytestmark = pytest.mark.gpu


def test_crops_match_reference_transform_image_bit_exact():
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_track.npz"))
    img = recipes.track_frame()
    frame = torch.from_numpy(img).cuda().unsqueeze(0)
    c = torch.from_numpy(g["centers"].astype(np.float32)).cuda()
    s = torch.from_numpy(g["scales"].astype(np.float32)).cuda()
    out = net_utils.crop_boxes_cuda(frame, c, s, None, recipes.TRACK_RES).cpu().numpy()
    np.testing.assert_array_equal(out, g["crops_u8"].astype(np.float32))
    fr32 = torch.from_numpy(img.astype(np.float32)).permute(2, 0, 1).contiguous().cuda().unsqueeze(0)
    out32 = net_utils.crop_boxes_cuda(fr32, c, s, None, recipes.TRACK_RES).cpu().numpy()
    np.testing.assert_allclose(out32, g["crops_f32"], rtol=0, atol=1e-4)
    rev = net_utils.crop_boxes_cuda(fr32, c, s, None, recipes.TRACK_RES, reverse_channels=True).cpu().numpy()
    np.testing.assert_array_equal(rev, out32[:, ::-1])


@pytest.mark.parametrize("res", [(256, 192), (64, 48)])
def test_crops_many_frames_vs_oracle(res):
    rng = np.random.default_rng(11)
    frames = rng.integers(0, 256, (3, 384, 512, 3), dtype=np.uint8)
    n = 24
    cx, cy = rng.uniform(-30, 540, n), rng.uniform(-30, 410, n)
    w, h = rng.uniform(20, 300, n), rng.uniform(20, 380, n)
    boxes = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], axis=1).astype(np.float32)     # float32 boxes, as the detector hands them over
    frame_of = rng.integers(0, 3, n).astype(np.int32)
    centers, scales = track_oracle.boxes_to_center_scale(boxes, res)
    c32, s32 = centers.astype(np.float32), scales.astype(np.float32)
    ref = np.stack([track_oracle.crop_boxes(frames[frame_of[i]], boxes[i:i + 1], res)[0] for i in range(n)])
    # the oracle computes centre/scale in double from the float32 boxes; the kernel receives float32 centre/scale
    ref32 = np.stack([track_oracle.warp_affine_fixed(frames[frame_of[i]], track_oracle.get_transform(c32[i].astype(np.float64), float(s32[i]), res)[:2], res)
                      .transpose(2, 0, 1) for i in range(n)]).astype(np.float32)
    out = net_utils.crop_boxes_cuda(torch.from_numpy(frames).cuda(), torch.from_numpy(c32).cuda(), torch.from_numpy(s32).cuda(),
                                    torch.from_numpy(frame_of).cuda(), res).cpu().numpy()
    np.testing.assert_array_equal(out, ref32)                       # same float32 centre/scale: bit-exact
    assert np.abs(out - ref).max() <= 1.0 or (np.abs(out - ref) > 1).mean() < 1e-3   # double centre/scale: at most a 1/32-px grid step apart
    assert out[:, :, 0, 0].shape == (n, 3)


def test_crops_empty_and_outside():
    frame = torch.full((1, 40, 50, 3), 200, dtype=torch.uint8, device="cuda")
    c = torch.tensor([[1000.0, 1000.0]], device="cuda")
    s = torch.tensor([64.0], device="cuda")
    out = net_utils.crop_boxes_cuda(frame, c, s, None, (64, 48))
    assert out.abs().max().item() == 0                              # box far outside the frame: zero border everywhere
    none = net_utils.crop_boxes_cuda(frame, c[:0], s[:0], None, (64, 48))
    assert tuple(none.shape) == (0, 3, 64, 48)


@pytest.mark.parametrize("n,thresh", [(1, 0.5), (37, 0.3), (300, 0.5), (1000, 0.7), (1024, 0.3)])
def test_nms_vs_oracle(n, thresh):
    rng = np.random.default_rng(n)
    F = 3
    xy = rng.uniform(0, 300, (F, n, 2)).astype(np.float32)
    wh = rng.uniform(5, 120, (F, n, 2)).astype(np.float32)
    sc = np.stack([rng.permutation(n) for _ in range(F)]).astype(np.float32)[..., None] / n
    dets = np.concatenate([xy, xy + wh, sc], axis=2)
    counts = np.array([n, max(1, n // 2), n], np.int32)
    for inclusive in (True, False):
        keep, num = net_utils.nms_cuda(torch.from_numpy(dets).cuda(), thresh, torch.from_numpy(counts).cuda(), inclusive)
        keep, num = keep.cpu().numpy(), num.cpu().numpy()
        for f in range(F):
            ref = track_oracle.nms_cpu(dets[f, :counts[f]], thresh, inclusive)
            assert num[f] == len(ref)
            np.testing.assert_array_equal(keep[f, :num[f]], ref)
            assert np.all(keep[f, num[f]:] == -1)


def test_nms_wrapper_signature_and_edge_cases():
    d = torch.tensor([[0, 0, 10, 10, .9], [1, 1, 11, 11, .8], [20, 20, 30, 30, .7], [0, 0, 10, 10, .95], [21, 21, 31, 31, .1]])
    assert net_utils.nms(d, 0.5).tolist() == [3, 2]                 # CPU tensor in -> cpu_nms rule, LongTensor out on the CPU
    assert net_utils.nms(d, 1.0).tolist() == [3, 1, 2, 4]
    assert net_utils.nms(d.cuda(), 1.0).tolist() == [3, 0, 1, 2, 4]  # CUDA tensor in -> the gpu rule (>)
    assert net_utils.nms(d[:0], 0.5).numel() == 0
    with pytest.raises(RuntimeError):
        net_utils.nms_cuda(torch.zeros(1, 2000, 5, device="cuda"), 0.5)     # more than 1024 boxes per frame


def test_pose_est_reference_signature_runs_on_gpu_crops():
    """pose_est(net, im, boxes): frame uploaded once, crops + PoseResNet + decode on the GPU == oracle crops + oracle net + oracle decode."""
    from flowtrack.pytorch_b200.pose import models as pose_models
    from flowtrack.pytorch_b200.tracking import synthetic
    from oracle import decode_oracle, pose_oracle
    torch.manual_seed(0)
    o_pose = synthetic.condition_pose_weights(pose_oracle.deconv("resnet50", 17, False), 0).eval()
    pose = pose_models.deconv("resnet50", 17, False)
    pose.load_state_dict(o_pose.state_dict(), strict=True)
    pose.cuda()
    rng = np.random.default_rng(5)
    im = rng.integers(0, 256, (240, 320, 3), dtype=np.uint8)
    boxes = np.array([[30.0, 20.0, 130.0, 220.0], [150.0, 40.0, 300.0, 200.0], [-10.0, 100.0, 80.0, 260.0]])
    res = (128, 96)
    kps = net_utils.pose_est(pose, im, boxes, inp_res=res)
    crops = track_oracle.crop_boxes(im, boxes, res)
    with torch.no_grad():
        hm = o_pose(torch.from_numpy(crops)).numpy()
    c, s = track_oracle.boxes_to_center_scale(boxes, res)
    xy, sc = decode_oracle.final_preds(hm, c, s, True)
    assert kps.shape == (3, 17, 3)
    assert np.abs(kps[:, :, :2] - xy).max() <= 1.0 * s.max() / res[0] + 1e-3      # <= 1 heat-map... image px per heat-map px scale
    assert np.abs(kps[:, :, 2:] - sc).max() < 2e-2


def test_flip_test_forward_matches_reference_recipe():
    """tools/pose/main.py:289-299 on the GPU (mirrored crops in the same batch) vs the oracle net + restated swaplr_image."""
    from flowtrack.pytorch_b200 import _lib, engine
    from flowtrack.pytorch_b200.pose import models as pose_models
    from flowtrack.pytorch_b200.pose.utils import evaluation
    from flowtrack.pytorch_b200.tracking import synthetic
    from oracle import pose_oracle
    torch.manual_seed(0)
    o_pose = synthetic.condition_pose_weights(pose_oracle.deconv("resnet50", 17, False), 0).eval()
    pose = pose_models.deconv("resnet50", 17, False)
    pose.load_state_dict(o_pose.state_dict(), strict=True)
    pose.cuda()
    x = torch.randn(3, 3, 128, 96, generator=torch.Generator().manual_seed(2))
    out = evaluation.flip_test_forward(pose, x.cuda()).cpu().numpy()
    with torch.no_grad():
        ref = track_oracle.flip_test_merge(o_pose(x).numpy(), o_pose(torch.flip(x, dims=[3])).numpy())
    assert out.shape == ref.shape == (3, 17, 32, 24)
    assert np.abs(out - ref).max() < 1e-3
    # the two kernels alone, bit-exact
    a = torch.randn(5, 7, 6, 9, device="cuda")
    f = torch.empty_like(a)
    _lib.check(_lib.lib().ft_flip_w(a.data_ptr(), f.data_ptr(), 5 * 7 * 6, 9, engine.stream_ptr()), "flip_w")
    assert torch.equal(f, torch.flip(a, dims=[3]))
    hm, hf = torch.randn(2, 17, 8, 6, device="cuda"), torch.randn(2, 17, 8, 6, device="cuda")
    from flowtrack.pytorch_b200.pose.utils import transforms
    perm = torch.tensor(transforms.swap_permutation(17), dtype=torch.int32, device="cuda")
    o = torch.empty_like(hm)
    _lib.check(_lib.lib().ft_flip_merge(hm.data_ptr(), hf.data_ptr(), perm.data_ptr(), 2, 17, 8, 6, o.data_ptr(), engine.stream_ptr()), "flip_merge")
    np.testing.assert_array_equal(o.cpu().numpy(), track_oracle.flip_test_merge(hm.cpu().numpy(), hf.cpu().numpy()))


def test_process_frame_entry_point(tmp_path):
    """tools/tracking/demo.py process_frame (reference :35-42) end to end on the GPU with a stand-in detector, and
    tools/flownet/demo.py writing a readable .flo file from a saved checkpoint."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("ft_tracking_demo", os.path.join(root, "tools", "tracking", "demo.py"))
    demo = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(demo)
    args = demo.parse_args(["--pose_backbone", "50", "--flow_net", "FlowNet2S"])
    torch.manual_seed(0)
    pose_net, flow_net = demo.build_nets(args)
    rng = np.random.default_rng(0)
    prev_im = rng.integers(0, 256, (128, 192, 3), dtype=np.uint8)
    cur_im = rng.integers(0, 256, (128, 192, 3), dtype=np.uint8)
    prev_dets = np.array([[20, 10, 90, 120, 0.9], [100, 30, 180, 110, 0.8]], np.float32)
    prev_kps = net_utils.pose_est(pose_net, prev_im, prev_dets[:, :4], inp_res=(128, 96))

    def fake_detector(im):          # im_detect contract: scores [R, classes], boxes [R, 4*classes]; person = class 1
        boxes = np.zeros((2, 8), np.float32)
        boxes[:, 4:] = [[22, 12, 92, 118], [5, 5, 40, 60]]
        return np.array([[0.1, 0.95], [0.5, 0.6]], np.float32), boxes

    dets, kps = demo.process_frame(fake_detector, pose_net, flow_net, prev_im, cur_im, prev_dets, prev_kps)
    assert dets.shape[1] == 5 and kps.shape == (dets.shape[0], 17, 3) and 2 <= dets.shape[0] <= 4
    assert np.all(np.diff(dets[:, 4]) <= 0) and np.isfinite(kps).all()
    # flow demo: checkpoint in, .flo out
    spec = importlib.util.spec_from_file_location("ft_flow_demo", os.path.join(root, "tools", "flownet", "demo.py"))
    fdemo = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fdemo)
    import cv2
    from flowtrack.pytorch_b200.flownet.model import models as flow_models
    from flowtrack.pytorch_b200.flownet.utils import flow_utils
    ck = tmp_path / "flow.pth.tar"
    torch.save({"arch": "FlowNet2S", "epoch": 1, "state_dict": flow_models.FlowNet2S(None).state_dict(), "best_EPE": 0.0}, str(ck))
    cv2.imwrite(str(tmp_path / "a.png"), prev_im)
    cv2.imwrite(str(tmp_path / "b.png"), cur_im)
    assert fdemo.main(["--input1", str(tmp_path / "a.png"), "--input2", str(tmp_path / "b.png"), "--resume", str(ck), "--save", str(tmp_path / "out")]) == 0
    flo = flow_utils.readFlow(str(tmp_path / "out" / "output.flo"))
    assert flo.shape == (128, 192, 2) and np.isfinite(flo).all()


def test_nms_oks_matches_reference_outputs():
    """ft_nms_oks == the reference's own nms_oks on the golden poses (tests/golden/reference_track.npz) and == the oracle on
    random pose sets, incl. the reference-signature wrapper."""
    from flowtrack.pytorch_b200.pose.utils import evaluation
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_track.npz"))
    preds = recipes.oks_poses()
    for th, kt in ((0.9, 0.0), (0.5, 0.2), (0.3, 0.5)):
        assert evaluation.nms_oks(preds, th, recipes.COCO_DELTA, kt) == g["oks_keep_%g_%g" % (th, kt)].tolist()
    for seed, n in ((1, 7), (2, 90), (3, 300)):
        p2 = recipes.oks_poses(n, seed)
        for i, p in enumerate(p2):                       # distinct scores for any n
            p["score"] = float(1.0 - i / (2.0 * n))
        assert evaluation.nms_oks(p2, 0.4, recipes.COCO_DELTA, 0.3) == track_oracle.nms_oks(p2, 0.4, recipes.COCO_DELTA, 0.3)
    assert evaluation.nms_oks([], 0.5, recipes.COCO_DELTA) == []


def test_crop_affine_s2d_equals_crop_then_pack():
    """ft_crop_affine_s2d (the tracker's fused path) == ft_crop_affine followed by ft_pack_image_s2d, bit for bit, from uint8 BGR
    frames addressed by strides (the current frame of every pair) and from float frames with reversed channels."""
    from flowtrack.pytorch_b200 import engine
    from flowtrack.pytorch_b200.tracking import net_utils
    rng = np.random.default_rng(11)
    F_, H, W, oh, ow, P = 3, 96, 128, 64, 48, 2
    pairs = torch.from_numpy(rng.integers(0, 255, (F_, 2, H, W, 3), dtype=np.uint8)).cuda()
    n = F_ * P
    centers = torch.tensor(rng.uniform(20, 100, (n, 2)), dtype=torch.float32).cuda()
    scales = torch.tensor(rng.uniform(40, 140, n), dtype=torch.float32).cuda()
    frame_of = (torch.arange(n, dtype=torch.int32) // P).cuda()
    L = _lib.lib()
    for frames, is_u8, rev in ((pairs[:, 1], 1, 0), (pairs[:, 1].permute(0, 3, 1, 2).float().contiguous(), 0, 1)):
        crops = net_utils.crop_boxes_cuda(frames, centers, scales, frame_of, (oh, ow), reverse_channels=bool(rev))
        want = engine.Act(n, oh // 2, ow // 2, 16, "cuda", xpad=(4, 4))
        _lib.check(L.ft_pack_image_s2d(crops.data_ptr(), n, oh, ow, want.struct(), 4, 4, gpu_util.sp()), "pack_s2d")
        got = engine.Act(n, oh // 2, ow // 2, 16, "cuda", xpad=(4, 4))
        if is_u8:
            sn, sh, sw, sc = frames.stride()
        else:
            sn, sc, sh, sw = frames.stride()
        _lib.check(L.ft_crop_affine_s2d(frames.data_ptr(), is_u8, sn, sc, sh, sw, H, W, 3, centers.data_ptr(), scales.data_ptr(), frame_of.data_ptr(), n, oh, ow,
                                        rev, got.struct(), 4, 4, gpu_util.sp()), "crop_s2d")
        torch.cuda.synchronize()
        assert torch.equal(got.hi, want.hi) and torch.equal(got.lo, want.lo)
        assert got.hi[:, :, :4].abs().max() == 0 and got.hi[:, :, -4:].abs().max() == 0          # the zero columns stay untouched


def test_pack_pair_u8_equals_float_pack():
    """ft_pack_pair_u8 (uint8 BGR frames) == ft_pack_pair on the same pair as float RGB [B,3,2,H,W]: the x_f32 output within float
    rounding of the mean (exact integer sums vs a float reduction), the split-fp16 view likewise; zero columns untouched."""
    from flowtrack.pytorch_b200 import engine
    rng = np.random.default_rng(5)
    B, H, W = 3, 64, 128
    u8 = torch.from_numpy(rng.integers(0, 255, (B, 2, H, W, 3), dtype=np.uint8)).cuda()
    pair = u8.flip(4).permute(0, 4, 1, 2, 3).float().contiguous()                               # [B,3(RGB),2,H,W]
    L = _lib.lib()
    xf_a, xf_b = torch.empty(B, 6, H, W, device="cuda"), torch.empty(B, 6, H, W, device="cuda")
    a, b = engine.Act(B, H, W, 6, "cuda", xpad=(4, 8)), engine.Act(B, H, W, 6, "cuda", xpad=(4, 8))
    ws = torch.zeros(L.ft_pack_pair_ws_floats(B), device="cuda")
    _lib.check(L.ft_pack_pair(pair.data_ptr(), B, H, W, 255.0, ws.data_ptr(), xf_a.data_ptr(), a.struct(), 4, 8, gpu_util.sp()), "pack_pair")
    sums = torch.zeros(B * 3, dtype=torch.int64, device="cuda")
    _lib.check(L.ft_pack_pair_u8(u8.data_ptr(), B, H, W, 1, 255.0, sums.data_ptr(), xf_b.data_ptr(), b.struct(), 4, 8, gpu_util.sp()), "pack_pair_u8")
    torch.cuda.synchronize()
    ref = (pair - pair.reshape(B, 3, -1).mean(-1).view(B, 3, 1, 1, 1)) / 255.0
    ref = torch.cat((ref[:, :, 0], ref[:, :, 1]), dim=1)
    assert (xf_b - ref).abs().max() < 1e-6 and (xf_a - xf_b).abs().max() < 1e-6
    assert (a.to_nchw(6) - b.to_nchw(6)).abs().max() < 1e-6 and (b.to_nchw(6) - ref).abs().max() < 1e-6
    assert b.hi[:, :, :4].abs().max() == 0 and b.hi[:, :, -8:].abs().max() == 0 and b.to_nchw()[:, 6:].abs().max() == 0
