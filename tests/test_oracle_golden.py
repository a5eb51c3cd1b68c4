"""CPU: the oracle restatement reproduces the fixtures generated from the REAL reference
(tests/golden/make_golden.py).  This is what pins the oracle (the reference has no
golden vectors of its own, only parameter-count doc-strings, which are checked here too)."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import decode_oracle, ops_oracle
from tests.golden import recipes


def _digest(sd):
    return hashlib.sha256("\n".join("%s %s" % (k, tuple(v.shape)) for k, v in sd.items()).encode()).hexdigest()


@pytest.mark.parametrize("depth", [50, 101, 152])
def test_pose_oracle_matches_reference(golden, depth):
    m = recipes.pose_oracle_model(depth)
    assert _digest(m.state_dict()) == str(golden["pose%d_keys" % depth])          # same state_dict layout
    assert sum(p.numel() for p in m.parameters()) == int(golden["pose%d_meta" % depth][4])
    with torch.no_grad():
        y = m(recipes.pose_input(depth)).numpy()
    ref = golden["pose%d_heatmap" % depth]
    assert y.shape == ref.shape
    assert np.abs(y - ref).max() <= 2e-4 * max(1.0, np.abs(ref).max())


def test_pose_param_counts():
    # measured from the reference factory
    from oracle import pose_oracle
    for depth, n in ((50, 33999697), (101, 52991825), (152, 68635473)):
        assert sum(p.numel() for p in pose_oracle.PoseResNetOracle(depth, 17).parameters()) == n


@pytest.mark.parametrize("name", recipes.FLOW_NAMES)
def test_flow_oracle_matches_reference(golden, name):
    m = recipes.flow_oracle_model(name)
    assert _digest(m.state_dict()) == str(golden["flow_%s_keys" % name])
    assert sum(p.numel() for p in m.parameters()) == int(golden["flow_%s_params" % name])
    with torch.no_grad():
        y = m(recipes.flow_input()).numpy()
    ref = golden["flow_%s_out" % name]
    assert y.shape == ref.shape
    assert np.abs(y - ref).max() <= 1e-3 * max(1.0, np.abs(ref).max()), np.abs(y - ref).max()


def test_flow_param_count_docstrings():
    # the reference's only known answers: models.py:17, FlowNetC.py:11, FlowNetSD.py:9, FlowNetFusion.py:9
    from oracle import flow_oracle as fo
    n = lambda m: sum(p.numel() for p in m.parameters())
    assert n(fo.FlowNet2(None)) == 162518834
    assert n(fo.FlowNet2C(None)) == 39175298
    assert n(fo.FlowNet2SD(None)) == 45371666
    assert n(fo.FlowNetFusion(None, batchNorm=False)) == 581226
    # FlowNetS.py:13 says 38,676,504; the 6-channel FlowNet2S instantiation measures 38,676,506 on the
    # real reference too -- the doc-string is off by two.
    assert n(fo.FlowNet2S(None)) == 38676506


def test_decode_oracle_matches_reference(golden):
    hm = recipes.decode_heatmaps().numpy()
    c, s = decode_oracle.final_preds(hm, golden["decode_center"], golden["decode_scale"], adjust_coords=True)
    c0, _ = decode_oracle.final_preds(hm, golden["decode_center"], golden["decode_scale"], adjust_coords=False)
    np.testing.assert_allclose(c, golden["decode_coords_adjust"], rtol=0, atol=1e-4)
    np.testing.assert_allclose(c0, golden["decode_coords_plain"], rtol=0, atol=1e-4)
    np.testing.assert_array_equal(s, golden["decode_scores"])
    assert np.all(c[0, 3] == decode_oracle.transform_preds(np.zeros((1, 1, 2), np.float32), golden["decode_center"][:1],
                                                           golden["decode_scale"][:1], (64, 48))[0, 0])
    np.testing.assert_allclose(decode_oracle.transform_preds(golden["tf_pts"], golden["decode_center"][:1],
                                                             golden["decode_scale"][:1], (64, 48)), golden["tf_out"], atol=1e-4)


def test_box_propagation_known_answer():
    # hand-computed: 2 people, 3 joints, constant flow (+2, -1); one joint masked out
    flow = np.zeros((2, 40, 60), np.float32)
    flow[0] += 2.0
    flow[1] -= 1.0
    kp = np.array([[[10, 10, 1], [20, 30, 1], [5, 5, 0]],
                   [[50, 35, .5], [58, 38, .9], [55, 20, .1]]], np.float32)
    b = decode_oracle.box_propagation(kp, flow)
    # person 0: shifted (12,9),(22,29) -> min (12,9) max (22,29), extend (0.75,1.5)
    np.testing.assert_allclose(b[0], [11.25, 7.5, 22.75, 30.5], atol=1e-6)
    # person 1: shifted (52,34),(60,37),(57,19) -> min (52,19) max (60,37); extend (.6,1.35); clip to W-1=59,H-1=39
    np.testing.assert_allclose(b[1], [51.4, 17.65, 59.0, 38.35], atol=1e-6)


def test_ops_c_restatement_vs_vectorised():
    g = torch.Generator().manual_seed(3)
    a, b = torch.randn(2, 40, 12, 16, generator=g), torch.randn(2, 40, 12, 16, generator=g)
    for (pad, k, md, s1, s2) in ((20, 1, 20, 1, 2), (4, 1, 4, 1, 1), (3, 3, 2, 2, 1)):
        oc = ops_oracle.correlation_c(a, b, pad, k, md, s1, s2)
        ot = ops_oracle.correlation_torch(a, b, pad, k, md, s1, s2)
        assert oc.shape == ot.shape
        assert (oc - ot).abs().max() < 5e-6
    img, flow = torch.randn(2, 3, 24, 32, generator=g), torch.randn(2, 2, 24, 32, generator=g) * 6
    rc, rt = ops_oracle.resample2d_c(img, flow), ops_oracle.resample2d_torch(img, flow)
    assert (rc - rt).abs().max() < 1e-5
    # Resample2d == grid_sample(bilinear, border, align_corners=True)
    yy, xx = torch.meshgrid(torch.arange(24.), torch.arange(32.), indexing="ij")
    grid = torch.stack(((xx + flow[:, 0]) / 31 * 2 - 1, (yy + flow[:, 1]) / 23 * 2 - 1), -1)
    gs = torch.nn.functional.grid_sample(img, grid, mode="bilinear", padding_mode="border", align_corners=True)
    assert (rc - gs).abs().max() < 1e-4
    x = torch.randn(2, 3, 24, 32, generator=g)
    assert (ops_oracle.channelnorm_c(x) - ops_oracle.channelnorm_torch(x)).abs().max() < 1e-6
    assert (ops_oracle.channelnorm_c(x) - x.norm(2, 1, keepdim=True)).abs().max() < 1e-6


def test_correlation_shape_flownetc():
    assert ops_oracle.correlation_shape(48, 64, 20, 1, 20, 1, 2) == (441, 48, 64)


# ---- tracking glue: person crops and NMS (row 2) ----
def test_crop_oracle_matches_reference_transform_image():
    """oracle.track_oracle.warp_affine_fixed (restated OpenCV bilinear warpAffine) == crops produced by the reference's
    own transform_image (tests/golden/reference_track.npz), bit for bit on the uint8 frame."""
    from oracle import track_oracle
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_track.npz"))
    img = recipes.track_frame()
    c, s = track_oracle.boxes_to_center_scale(recipes.TRACK_BOXES, recipes.TRACK_RES)
    np.testing.assert_array_equal(c, g["centers"])
    np.testing.assert_array_equal(s, g["scales"])
    crops = track_oracle.crop_boxes(img, recipes.TRACK_BOXES, recipes.TRACK_RES)
    np.testing.assert_array_equal(crops.astype(np.uint8), g["crops_u8"])
    crops_f = track_oracle.crop_boxes(img.astype(np.float32), recipes.TRACK_BOXES, recipes.TRACK_RES)
    np.testing.assert_allclose(crops_f, g["crops_f32"], rtol=0, atol=1e-4)


def test_crop_oracle_matches_cv2_on_random_boxes():
    cv2 = pytest.importorskip("cv2")
    from oracle import track_oracle
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (120, 160, 3), dtype=np.uint8)
    for i in range(12):
        c = (rng.uniform(-20, 180), rng.uniform(-20, 140))
        s = rng.uniform(15, 300)
        res = (64, 48) if i % 2 else (128, 96)
        M = track_oracle.get_transform(c, s, res)[:2]
        np.testing.assert_array_equal(track_oracle.warp_affine_fixed(img, M, res), cv2.warpAffine(img, M, (res[1], res[0])))


def test_nms_oracle_hand_cases_and_properties():
    from oracle import track_oracle
    d = np.array([[0, 0, 10, 10, .9], [1, 1, 11, 11, .8], [20, 20, 30, 30, .7], [0, 0, 10, 10, .95], [21, 21, 31, 31, .1]], np.float32)
    # IoU(box0, box1) = 100/142 = 0.704, IoU(box2, box4) = 0.704, box3 == box0
    assert track_oracle.nms_cpu(d, 0.5).tolist() == [3, 2]
    assert track_oracle.nms_cpu(d, 0.75).tolist() == [3, 1, 2, 4]
    assert track_oracle.nms_cpu(d, 1.0, inclusive=True).tolist() == [3, 1, 2, 4]      # identical boxes: IoU == 1 >= 1 suppresses
    assert track_oracle.nms_cpu(d, 1.0, inclusive=False).tolist() == [3, 0, 1, 2, 4]  # the GPU rule (>) keeps them
    rng = np.random.default_rng(0)
    xy = rng.uniform(0, 200, (200, 2)).astype(np.float32)
    wh = rng.uniform(5, 80, (200, 2)).astype(np.float32)
    dets = np.concatenate([xy, xy + wh, rng.permutation(200).astype(np.float32)[:, None] / 200], axis=1)
    keep = track_oracle.nms_cpu(dets, 0.3)
    assert len(set(keep.tolist())) == len(keep) and np.all(np.diff(dets[keep, 4]) < 0)   # unique, descending score
    assert keep[0] == dets[:, 4].argmax()
    again = track_oracle.nms_cpu(dets[keep], 0.3)                                         # idempotent on its own output
    assert again.tolist() == list(range(len(keep)))


def test_flip_oracle_matches_reference_swaplr_image():
    from oracle import track_oracle
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_track.npz"))
    hm = recipes.flip_heatmaps()
    for ds, k in (("coco", 17), ("mpii", 16), ("aic", 14)):
        np.testing.assert_array_equal(track_oracle.swaplr_image(hm[:, :k], ds), g["swap_" + ds])
    from flowtrack.pytorch_b200.pose.utils import transforms
    for ds, k in (("coco", 17), ("mpii", 16), ("aic", 14)):
        perm = transforms.swap_permutation(k, ds)
        np.testing.assert_array_equal(hm[:, :k][:, perm][:, :, :, ::-1], g["swap_" + ds])     # the permutation form the kernel uses


def test_oks_oracle_matches_reference_nms_oks():
    from oracle import track_oracle
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_track.npz"))
    preds = recipes.oks_poses()
    for th, kt in ((0.9, 0.0), (0.5, 0.2), (0.3, 0.5)):
        assert track_oracle.nms_oks(preds, th, recipes.COCO_DELTA, kt) == g["oks_keep_%g_%g" % (th, kt)].tolist()
    j = np.array([p["joints"] for p in preds])
    oks = track_oracle.compute_oks(j[1:], j[0], np.full(len(preds) - 1, 5000.0), recipes.COCO_DELTA, ground_truth=False, threshold=0.2)
    np.testing.assert_allclose(oks, g["oks_pairwise0"], rtol=1e-12, atol=0)


def test_nms_oracle_matches_reference_cpu_nms():
    """oracle.track_oracle.nms_cpu == the reference's own cpu_nms (nms.c compiled unmodified against oracle/ref_shims): against the
    committed fixtures everywhere, and against the library itself where it was built (this container)."""
    from oracle import track_oracle
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_track.npz"))
    for n, seed, th in recipes.NMS_CASES:
        dets = recipes.nms_boxes(n, seed)
        mine = track_oracle.nms_cpu(dets, th, inclusive=True)
        np.testing.assert_array_equal(mine, g["nms_keep_%d_%d_%g" % (n, seed, th)])
        live = track_oracle.ref_cpu_nms(dets, th)
        if live is not None:
            np.testing.assert_array_equal(mine, live)
    if track_oracle.ref_cpu_nms(recipes.nms_boxes(5, 9), 0.5) is not None:       # a few more thresholds / sizes against the live library
        for n, seed, th in [(64, 11, 0.1), (257, 12, 0.45), (700, 13, 0.9)]:
            dets = recipes.nms_boxes(n, seed)
            np.testing.assert_array_equal(track_oracle.nms_cpu(dets, th), track_oracle.ref_cpu_nms(dets, th))


def test_flowvis_oracle_matches_reference_flowlib():
    """oracle/flowvis_oracle.py == the reference's own flowlib.flow_to_image / make_color_wheel (fixtures by make_golden.py)."""
    import os
    from oracle import flowvis_oracle
    fix = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_flowvis.npz"))
    assert np.array_equal(flowvis_oracle.make_color_wheel(), fix["flowvis_wheel"])
    for i, flow in enumerate(recipes.flowvis_flows()):
        assert np.array_equal(flowvis_oracle.flow_to_image(flow), fix["flowvis_img%d" % i])
