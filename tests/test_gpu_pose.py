"""GPU parity of the PoseResNet drop-in vs the CPU oracle (conditioned synthetic weights).
Gate (BASELINE.json north_star): heat-maps within 1e-3 max-abs, decoded key-points within 1 px."""
import numpy as np
import pytest
import torch

from flowtrack.pytorch_b200 import _lib
from flowtrack.pytorch_b200.pose import models
from oracle import decode_oracle
from tests.golden import recipes

pytestmark = pytest.mark.gpu


def _pair(depth, calib_hw=None):
    orc = recipes.pose_oracle_model(depth, calib_hw=calib_hw)
    m = models.deconv("resnet%d" % depth, 17, False)
    m.load_state_dict(orc.state_dict(), strict=True)
    return orc, m.cuda()


@pytest.mark.parametrize("depth", [50, 101, 152])
def test_pose_matches_golden_fixture(golden, depth):
    orc, m = _pair(depth)
    out = m(recipes.pose_input(depth).cuda()).cpu()
    ref = torch.from_numpy(golden["pose%d_heatmap" % depth])      # produced by the REAL reference modules
    assert out.shape == ref.shape
    assert (out - ref).abs().max() < 1e-3, (out - ref).abs().max()


def test_pose_r50_256x192_vs_oracle_split_and_fast():
    orc, m = _pair(50, calib_hw=(256, 192))
    x = torch.randn(4, 3, 256, 192, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        ref = orc(x)
    assert ref.abs().max() > 1.0                                   # conditioned weights: the gate is not vacuous
    out = m(x.cuda()).cpu()
    err = (out - ref).abs().max().item()
    print("pose r50 256x192 split max-abs %.3g (heat-map abs-max %.2f)" % (err, ref.abs().max().item()))
    assert out.shape == (4, 17, 64, 48)
    assert err < 1e-3, err
    # decoded key-points within 1 px (heat-map pixels -> image pixels through the same crop transform)
    center = np.tile(np.array([[96.0, 128.0]]), (4, 1))
    scale = np.full(4, 256.0)
    kr, _ = decode_oracle.final_preds(ref.numpy(), center, scale, True)
    ko, _ = decode_oracle.final_preds(out.numpy(), center, scale, True)
    assert np.abs(kr - ko).max() <= 1.0
    # the SIMT cross-check implementation agrees with the tensor-core one
    m._impl = _lib.IMPL_SIMT
    sim = m(x.cuda()).cpu()
    m._impl = _lib.IMPL_TCGEN05
    assert (sim - ref).abs().max() < 1e-3 and (sim - out).abs().max() < 1e-3
    # fast mode (single fp16 pass) is reported, not gated at 1e-3
    m.precision = "fast"
    fast = m(x.cuda()).cpu()
    m.precision = "split"
    ferr = (fast - ref).abs().max().item()
    print("pose r50 256x192: split max-abs %.3g, fast max-abs %.3g" % (err, ferr))
    assert ferr < 0.5


def test_pose_graph_replay_is_deterministic_and_tracks_new_inputs():
    orc, m = _pair(50)
    x1, x2 = recipes.pose_input(50).cuda(), recipes.pose_input(50).cuda() * 0.5
    a1 = m(x1)
    a2 = m(x2)
    b1 = m(x1)
    assert torch.equal(a1, b1) and not torch.equal(a1, a2)
    m.use_cuda_graph = False
    m.refresh()
    c1 = m(x1)
    assert torch.equal(a1, c1)


def test_pose_forward_with_the_7x7_stem_form(monkeypatch):
    """FT_STEM_S2D=0 keeps the original 7x7 FT_K_ROW stem; both stem forms give the same heat-maps."""
    from flowtrack.pytorch_b200.pose import models as pose_models
    from flowtrack.pytorch_b200.tracking import synthetic
    torch.manual_seed(0)
    m = synthetic.condition_pose_weights(pose_models.deconv("resnet50", 17, False), 0)
    x = torch.randn(2, 3, 128, 96, generator=torch.Generator().manual_seed(1)).cuda()
    a = m.cuda()(x).clone()
    monkeypatch.setenv("FT_STEM_S2D", "0")
    m.refresh()
    b = m(x).clone()
    assert m._layers["stem"].cin_p == 8
    assert (a - b).abs().max() < 2e-4


def test_pose_r50_384x288_vs_oracle():
    """BASELINE config 5 crop size (384x288 -> 96x72 heat-maps), batch not a multiple of the tile's image count."""
    orc, m = _pair(50, calib_hw=(384, 288))
    x = torch.randn(3, 3, 384, 288, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        ref = orc(x)
    out = m(x.cuda()).cpu()
    err = (out - ref).abs().max().item()
    print("pose r50 384x288 split max-abs %.3g (heat-map abs-max %.2f)" % (err, ref.abs().max().item()))
    assert out.shape == (3, 17, 96, 72)
    assert err < 1e-3, err
