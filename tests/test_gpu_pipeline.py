"""GPU: the joint frame-pair step (flow -> box propagation, pose -> decode) through FramePairPipeline vs the CPU oracle,
and the reference-signature glue (pose_est / flow_est)."""
import numpy as np
import pytest
import torch

from flowtrack.pytorch_b200.flownet.model import models as flow_models
from flowtrack.pytorch_b200.pose import models as pose_models
from flowtrack.pytorch_b200.tracking import net_utils
from flowtrack.pytorch_b200.tracking import pipeline as pl
from flowtrack.pytorch_b200.tracking import synthetic
from oracle import decode_oracle, flow_oracle, pose_oracle

pytestmark = pytest.mark.gpu


def _nets():
    torch.manual_seed(0)
    o_pose = synthetic.condition_pose_weights(pose_oracle.deconv("resnet50", 17, False), 0).eval()
    o_flow = flow_oracle.FlowNet2S(None).eval()
    pose = pose_models.deconv("resnet50", 17, False)
    pose.load_state_dict(o_pose.state_dict(), strict=True)
    flow = flow_models.FlowNet2S(None)
    flow.load_state_dict(o_flow.state_dict(), strict=True)
    return o_pose, o_flow, pose.cuda(), flow.cuda()


@pytest.mark.parametrize("graph", [True, False])
def test_frame_pair_pipeline_matches_oracle(graph):
    o_pose, o_flow, pose, flow = _nets()
    pairs, people, frame_hw, crop_hw = 3, 2, (128, 192), (128, 96)
    pipe = pl.FramePairPipeline(pose, flow, pairs, people, frame_hw, crop_hw, "cuda:0", use_cuda_graph=graph)
    d = synthetic.frame_pair_batch(pairs, people, frame_hw, crop_hw, seed=3)
    for _ in range(2):                                                       # second call replays the graph / reuses the plan
        kps, boxes = pipe.process(d["pair"], d["crops"], d["centers"], d["scales"], d["prev_keypoints"])
    kps, boxes = kps.clone().numpy(), boxes.clone().numpy()
    with torch.no_grad():
        r_flow, r_hm = o_flow(d["pair"]), o_pose(d["crops"])
    assert (pipe.heatmaps.cpu() - r_hm).abs().max() < 1e-3                   # north-star gate on heat-maps
    assert (pipe.flow.cpu() - r_flow).abs().max() < 1e-3 * max(1.0, r_flow.abs().max().item())
    r_xy, r_sc = decode_oracle.final_preds(r_hm.numpy(), d["centers"].numpy(), d["scales"].numpy(), True)
    r_boxes = np.concatenate([decode_oracle.box_propagation(d["prev_keypoints"][i * people:(i + 1) * people].numpy(), r_flow[i].numpy())
                              for i in range(pairs)])
    assert np.abs(kps[:, :, :2] - r_xy).max() <= 1.0                         # key-points within 1 px
    assert np.abs(kps[:, :, 2:] - r_sc).max() < 1e-3
    assert np.abs(boxes - r_boxes).max() <= 1.0
    assert pipe.launches_per_step > 0 and pipe.h2d_bytes > 0 and pipe.d2h_bytes == 4 * pairs * people * (17 * 3 + 4)


def test_frame_pair_pipeline_crops_cut_on_the_device():
    """crops_from_frames: the crops are warpAffine'd out of the current frame on the GPU (raw 0..255 BGR values, as the
    reference's pose_est feeds them) -- compared with the oracle crop + oracle nets on the same frames."""
    from oracle import track_oracle
    o_pose, o_flow, pose, flow = _nets()
    pairs, people, frame_hw, crop_hw = 2, 2, (128, 192), (128, 96)
    pipe = pl.FramePairPipeline(pose, flow, pairs, people, frame_hw, crop_hw, "cuda:0", crops_from_frames=True)
    d = synthetic.frame_pair_batch(pairs, people, frame_hw, crop_hw, seed=4)
    d["pair"] = d["pair"].round()                                           # integer pixel values: the float crop arithmetic is exact
    for _ in range(2):
        kps, boxes = pipe.process(d["pair"], None, d["centers"], d["scales"], d["prev_keypoints"])
    kps = kps.clone().numpy()
    n = pairs * people
    crops = np.stack([track_oracle.warp_affine_fixed(d["pair"][i // people, :, 1].permute(1, 2, 0).numpy()[:, :, ::-1].copy(),
                                                     track_oracle.get_transform(d["centers"][i].double().numpy(), float(d["scales"][i]), crop_hw)[:2], crop_hw)
                      .transpose(2, 0, 1) for i in range(n)])
    assert np.array_equal(pipe.crops.cpu().numpy(), crops)                  # bit-exact crops
    with torch.no_grad():
        r_hm = o_pose(torch.from_numpy(crops))
    amax = r_hm.abs().max().item()
    assert (pipe.heatmaps.cpu() - r_hm).abs().max() < 1e-3 * max(1.0, amax / 4.0)      # 1e-3 at the calibrated heat-map scale (abs-max ~4)
    r_xy, r_sc = decode_oracle.final_preds(r_hm.numpy(), d["centers"].numpy(), d["scales"].numpy(), True)
    assert np.abs(kps[:, :, :2] - r_xy).max() <= 1.0
    assert pipe.h2d_bytes == 4 * (d["pair"].numel() + n * (3 + 17 * 3))     # no crops cross PCIe


def test_pose_est_and_flow_est_reference_signatures():
    o_pose, o_flow, pose, flow = _nets()
    rng = np.random.RandomState(0)
    im0 = rng.randint(0, 255, (128, 192, 3)).astype(np.uint8)
    im1 = rng.randint(0, 255, (128, 192, 3)).astype(np.uint8)
    boxes = np.array([[20., 10., 90., 120.], [100., 30., 180., 110.]])
    kps = net_utils.pose_est(pose, im1, boxes, inp_res=(128, 96))
    assert kps.shape == (2, 17, 3)
    f = net_utils.flow_est(flow, im0, im1)
    assert f.shape == (2, 128, 192)
    # oracle: same crops through the reference-equivalent modules
    centers, scales = net_utils.boxes_to_center_scale(boxes, (128, 96))
    crops = torch.stack([net_utils.transform_image(im1, c, s, (128, 96)) for c, s in zip(centers, scales)]).float()
    with torch.no_grad():
        r_hm = o_pose(crops)
        pair = torch.from_numpy(np.stack([im0[..., ::-1].copy(), im1[..., ::-1].copy()])[None].transpose(0, 4, 1, 2, 3).astype(np.float32))
        r_f = o_flow(pair)[0].numpy()
    r_xy, r_sc = decode_oracle.final_preds(r_hm.numpy(), centers, scales, True)
    assert np.abs(kps[:, :, :2] - r_xy).max() <= 1.0
    assert np.abs(f - r_f).max() < 1e-3 * max(1.0, np.abs(r_f).max())
