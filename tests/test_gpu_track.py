"""GPU parity: person crops (ft_crop_affine) and box NMS (ft_nms) vs the CPU oracle / the reference-generated crops."""
import os

import numpy as np
import pytest
import torch

from flowtrack.pytorch_b200.tracking import net_utils
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
