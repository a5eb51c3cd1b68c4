"""GPU parity of the FlowNet2 family drop-ins vs the fixtures produced by the REAL reference modules and vs
the CPU oracle.  Tolerance: gpu_util.FLOW_TOL = 1e-3 x max|flow| max-abs in split mode for every variant (flows here are
tens of pixels), fast mode reported only."""
import numpy as np
import pytest
import torch

from flowtrack.pytorch_b200 import _lib
from flowtrack.pytorch_b200.flownet.model import models
from tests import gpu_util
from tests.golden import recipes

pytestmark = pytest.mark.gpu


def _pair(name):
    orc = recipes.flow_oracle_model(name)
    m = getattr(models, name)(None)
    m.load_state_dict(orc.state_dict(), strict=True)
    return orc, m.cuda()


@pytest.mark.parametrize("name", recipes.FLOW_NAMES)
def test_flow_matches_golden_fixture(golden, name):
    orc, m = _pair(name)
    out = m(recipes.flow_input().cuda()).cpu()
    ref = torch.from_numpy(golden["flow_%s_out" % name])
    assert out.shape == ref.shape
    err, mag = (out - ref).abs().max().item(), ref.abs().max().item()
    print("%s: max-abs %.3g on flows up to %.1f px" % (name, err, mag))
    assert err <= gpu_util.flow_tol(ref), (err, mag)


def test_flownet2s_384x512_batch_vs_oracle():
    orc, m = _pair("FlowNet2S")
    x = recipes.flow_input(2, (384, 512), seed=3)
    with torch.no_grad():
        ref = orc(x)
    out = m(x.cuda()).cpu()
    err, mag = (out - ref).abs().max().item(), ref.abs().max().item()
    assert out.shape == (2, 2, 384, 512)
    assert err <= gpu_util.flow_tol(ref), (err, mag)
    m._impl = _lib.IMPL_SIMT
    sim = m(x.cuda()).cpu()
    m._impl = _lib.IMPL_TCGEN05
    assert (sim - ref).abs().max() <= gpu_util.flow_tol(ref)
    m.precision = "fast"
    fast = m(x.cuda()).cpu()
    print("FlowNet2S 384x512: split max-abs %.3g, fast max-abs %.3g, |flow| max %.1f" % (err, (fast - ref).abs().max().item(), mag))


def test_native_op_modules_keep_reference_signatures():
    from flowtrack.pytorch_b200.flownet.networks.channelnorm_package.modules.channelnorm import ChannelNorm
    from flowtrack.pytorch_b200.flownet.networks.correlation_package.modules.correlation import Correlation
    from flowtrack.pytorch_b200.flownet.networks.resample2d_package.modules.resample2d import Resample2d
    from oracle import ops_oracle
    g = torch.Generator().manual_seed(0)
    a, b = torch.randn(1, 32, 16, 24, generator=g), torch.randn(1, 32, 16, 24, generator=g)
    c = Correlation(pad_size=20, kernel_size=1, max_displacement=20, stride1=1, stride2=2, corr_multiply=1)
    assert (c(a.cuda(), b.cuda()).cpu() - ops_oracle.correlation_torch(a, b)).abs().max() < 1e-5
    img, fl = torch.randn(1, 3, 16, 24, generator=g), torch.randn(1, 2, 16, 24, generator=g) * 3
    assert (Resample2d()(img.cuda(), fl.cuda()).cpu() - ops_oracle.resample2d_torch(img, fl)).abs().max() < 1e-5
    assert (ChannelNorm()(img.cuda()).cpu() - ops_oracle.channelnorm_torch(img)).abs().max() < 1e-6
    with pytest.raises(RuntimeError):
        c(a, b)                                            # CPU tensors: no fallback
