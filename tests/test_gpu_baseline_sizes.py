"""GPU parity at the BASELINE.json configuration sizes (not reduced stand-ins), against the CPU oracle on the same seeded
inputs.  One tolerance everywhere (tests/gpu_util.py): heat-maps 1e-3 max-abs, flows 1e-3 x max|flow|, key-points 1 px.

  cfg 2  PoseResNet-50   [64,3,256,192]                      (lib/pose/models/pose_deconv.py:32-46)
  cfg 3  FlowNet2-S      [32,3,2,384,512]                    (lib/flownet/model/models.py:254-292)
  cfg 3' FlowNet2-C      [2,3,2,384,512]: the tcgen05 correlation runs in-network at its 48x64 geometry (FlowNetC.py:71-128)
  cfg 4  joint pipeline  16 pairs + 64 crops, serial tracking chain (tools/tracking/demo.py:35-42)
  cfg 5  PoseResNet-152  [4,3,384,288] + full FlowNet2 [1,3,2,384,512]   (resnet.py:51-55, models.py:109-178)
The oracle forwards run on the host cores (a few seconds each at these sizes)."""
import numpy as np
import pytest
import torch

from flowtrack.pytorch_b200.flownet.model import models as flow_models
from flowtrack.pytorch_b200.pose import models as pose_models
from oracle import decode_oracle
from tests import gpu_util
from tests.golden import recipes

pytestmark = pytest.mark.gpu


def _pose_pair(depth, hw):
    orc = recipes.pose_oracle_model(depth, calib_hw=hw, calib_batch=2)
    m = pose_models.deconv("resnet%d" % depth, 17, False)
    m.load_state_dict(orc.state_dict(), strict=True)
    return orc, m.cuda()


def _flow_pair(name):
    orc = recipes.flow_oracle_model(name)
    m = getattr(flow_models, name)(None)
    m.load_state_dict(orc.state_dict(), strict=True)
    return orc, m.cuda()


def _check_pose(depth, batch, hw, seed):
    orc, m = _pose_pair(depth, hw)
    x = torch.randn(batch, 3, hw[0], hw[1], generator=torch.Generator().manual_seed(seed))
    with torch.no_grad():
        ref = orc(x)
    out = m(x.cuda()).cpu()
    err, amax = (out - ref).abs().max().item(), ref.abs().max().item()
    print("PoseResNet-%d %s: heat-map max-abs %.3g (abs-max %.2f)" % (depth, tuple(x.shape), err, amax))
    assert out.shape == (batch, 17, hw[0] // 4, hw[1] // 4) and amax > 1.0
    assert err < gpu_util.HEATMAP_TOL, err
    center = np.tile(np.array([[hw[1] / 2.0, hw[0] / 2.0]]), (batch, 1))
    scale = np.full(batch, float(hw[0]))
    kr, _ = decode_oracle.final_preds(ref.numpy(), center, scale, True)
    ko, _ = decode_oracle.final_preds(out.numpy(), center, scale, True)
    assert np.abs(kr - ko).max() <= gpu_util.KEYPOINT_TOL_PX


def _check_flow(name, batch, seed):
    orc, m = _flow_pair(name)
    x = recipes.flow_input(batch, (384, 512), seed=seed)
    with torch.no_grad():
        ref = orc(x)
    out = m(x.cuda()).cpu()
    err, mag = (out - ref).abs().max().item(), ref.abs().max().item()
    print("%s %s: flow max-abs %.3g on flows up to %.1f px" % (name, tuple(x.shape), err, mag))
    assert out.shape == (batch, 2, 384, 512)
    assert err <= gpu_util.flow_tol(ref), (err, mag)
    return m


def test_cfg2_pose_resnet50_batch64_256x192():
    _check_pose(50, 64, (256, 192), seed=20)


def test_cfg3_flownet2s_batch32_384x512():
    _check_flow("FlowNet2S", 32, seed=21)


def test_cfg3_flownet2c_384x512_runs_the_tensor_core_correlation():
    m = _check_flow("FlowNet2C", 2, seed=22)
    plan = next(iter(m._plans.values()))[0]
    assert any(c[0] == "correlation" and c[1] is not None for c in plan.calls)
    assert tuple(plan.taps["corr"].shape) == (2, 441, 4, 24, 32)          # the parity-planar output of corr_pair_kernel at 48x64


def test_cfg5_pose_resnet152_batch4_384x288():
    _check_pose(152, 4, (384, 288), seed=23)


def test_cfg5_flownet2_full_stack_384x512():
    _check_flow("FlowNet2", 1, seed=24)


def test_cfg5_flownet2_autotuned_plan_with_concurrent_branches(monkeypatch):
    """The production configuration of BASELINE config 5's flow network: per-layer autotuning on (split-K is chosen for most
    layers at one pair) and the plan's branches (FlowNetSD next to the C -> S -> S chain, flow heads next to the deconvolutions)
    running as parallel graph paths -- each branch owns its split-K workspace.  Same oracle gate, and stable over replays."""
    monkeypatch.setenv("FT_AUTOTUNE", "1")
    m = _check_flow("FlowNet2", 1, seed=24)
    plan = next(iter(m._plans.values()))[0]
    assert plan._ntags > 5 and len(plan.tuned) > 60
    x = recipes.flow_input(1, (384, 512), seed=24).cuda()
    a = m(x).clone()
    for _ in range(5):
        assert torch.equal(m(x), a)
