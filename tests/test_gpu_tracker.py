"""GPU: the device-resident tracking step (tracking/tracker.py: flow -> box propagation -> merge + NMS -> crops -> pose ->
decode, the serial chain of tools/tracking/demo.py:35-42) against the CPU oracle restatement of process_frame, step by step
over a short synthetic video, with and without CUDA graphs, with detector boxes merged in, and with a person dropping out."""
import numpy as np
import pytest
import torch

from flowtrack.pytorch_b200.flownet.model import models as flow_models
from flowtrack.pytorch_b200.pose import models as pose_models
from flowtrack.pytorch_b200.tracking import synthetic
from flowtrack.pytorch_b200.tracking import tracker as tk
from oracle import flow_oracle, pose_oracle, track_oracle
from tests import gpu_util

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


def _video(F, T, hw, seed):
    """T+1 frames per stream: smooth BGR uint8 textures that drift a few pixels per frame."""
    rng = np.random.default_rng(seed)
    H, W = hw
    base = rng.integers(0, 255, (F, H + 32, W + 32, 3), dtype=np.uint8)
    return np.stack([np.stack([base[f, 2 * t:2 * t + H, 3 * t:3 * t + W] for t in range(T + 1)]) for f in range(F)])      # [F,T+1,H,W,3]


def _check_step(o_pose, o_flow, frames_t, prev, got, det, det_count, crop_hw, P):
    p_dets, p_count, p_kps = prev
    g_dets, g_count, g_kps = got
    F = frames_t.shape[0]
    K = p_kps.shape[1]
    for f in range(F):
        n = int(p_count[f])
        d = det[f, :int(det_count[f])] if det is not None else None
        w_dets, w_kps, _ = track_oracle.process_frame(o_pose, o_flow, frames_t[f, 0], frames_t[f, 1], p_dets[f, :n], p_kps.reshape(F, P, K, 3)[f, :n], d,
                                                      inp_res=crop_hw, thresh=0.3, max_people=P)
        m = w_dets.shape[0]
        assert int(g_count[f]) == m, (f, int(g_count[f]), m)
        assert np.abs(g_dets[f, :m] - w_dets).max() <= gpu_util.KEYPOINT_TOL_PX          # boxes ride on flow look-ups
        assert np.array_equal(g_dets[f, :m, 4], w_dets[:, 4])                            # scores are carried over exactly
        gk = g_kps.reshape(F, P, K, 3)[f]
        assert np.abs(gk[:m, :, :2] - w_kps[:, :, :2]).max() <= gpu_util.KEYPOINT_TOL_PX
        assert np.abs(gk[:m, :, 2] - w_kps[:, :, 2]).max() < gpu_util.HEATMAP_TOL * 4
        assert np.all(gk[m:] == 0) and np.all(g_dets[f, m:] == 0)                        # empty slots are zeroed


@pytest.mark.parametrize("graph", [True, False])
def test_tracking_pipeline_matches_process_frame_oracle(graph):
    o_pose, o_flow, pose, flow = _nets()
    F, P, T, D = 2, 3, 3, 2
    frame_hw, crop_hw = (128, 192), (128, 96)
    video = _video(F, T, frame_hw, seed=5)
    pipe = tk.TrackingPipeline(pose, flow, F, P, frame_hw, crop_hw, "cuda:0", max_det=D, use_cuda_graph=graph)
    # initial state: P-1 / P people per stream with boxes on a grid and key-points inside them
    d0 = synthetic.frame_pair_batch(F, P, frame_hw, crop_hw, seed=9)
    dets = np.concatenate([d0["boxes"].numpy().reshape(F, P, 4), np.linspace(0.9, 0.5, F * P, dtype=np.float32).reshape(F, P, 1)], axis=2)
    count = np.array([P - 1, P], np.int32)
    dets[0, P - 1] = 0
    kps = d0["prev_keypoints"].numpy().copy()
    kps.reshape(F, P, 17, 3)[0, P - 1] = 0
    pipe.seed(dets, count, kps)
    state = (dets.copy(), count.copy(), kps.copy())
    rng = np.random.default_rng(3)
    for t in range(T):
        frames_t = np.ascontiguousarray(video[:, t:t + 2])                                # [F,2,H,W,3] (prev, cur)
        det = np.zeros((F, D, 5), np.float32)
        det_count = np.zeros(F, np.int32)
        if t == 1:                                                                        # a detector contributes at step 1
            det[0, 0] = (20, 10, 80, 110, 0.95)                                           # a new person
            det[0, 1] = np.append(state[0][0, 0, :4] + 2.0, 0.99)                         # overlaps person 0: the higher score wins the NMS
            det_count[:] = (2, 0)
        got = tuple(x.clone().numpy() if torch.is_tensor(x) else x for x in pipe.process(torch.from_numpy(frames_t), det, det_count))
        _check_step(o_pose, o_flow, frames_t, state, got, det, det_count, crop_hw, P)
        state = got                                                                       # teacher forcing: the oracle follows the device state
    assert pipe.launches_per_step > 0 and pipe.h2d_bytes >= frames_t.size and pipe.d2h_bytes == 4 * (F * P * 5 + F + F * P * 17 * 3)


def test_tracking_pipeline_two_lanes_in_flight_equals_synchronous_steps():
    """submit() on both lanes before collect(): flow(t+1) overlaps tail(t); results equal the one-step-at-a-time run."""
    o_pose, o_flow, pose, flow = _nets()
    F, P, T = 2, 2, 4
    frame_hw, crop_hw = (128, 192), (128, 96)
    video = torch.from_numpy(_video(F, T, frame_hw, seed=6))
    d0 = synthetic.frame_pair_batch(F, P, frame_hw, crop_hw, seed=2)
    dets = np.concatenate([d0["boxes"].numpy().reshape(F, P, 4), np.full((F, P, 1), 0.8, np.float32)], axis=2)
    outs = []
    for pipelined in (False, True):
        pipe = tk.TrackingPipeline(pose, flow, F, P, frame_hw, crop_hw, "cuda:0")
        pipe.seed(dets, np.full(F, P, np.int32), d0["prev_keypoints"])
        res, pending = [], []
        for t in range(T):
            l = pipe.submit(pipe.stage(video[:, t:t + 2].contiguous()))
            pending.append(l)
            if not pipelined or len(pending) > 1:
                res.append(tuple(x.clone() for x in pipe.collect(pending.pop(0))))
        while pending:
            res.append(tuple(x.clone() for x in pipe.collect(pending.pop(0))))
        outs.append(res)
    for a, b in zip(*outs):
        for x, y in zip(a, b):
            assert torch.equal(x, y)
