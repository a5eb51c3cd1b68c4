"""GPU: the reference's OWN Python (lib/flownet/model/models.py, networks/*.py, */modules/*.py, */functions/*.py, staged
unmodified under baseline/_ref by tools/stage_reference.py) running over the three `_ext/<op>/__init__.py` drop-in shims this
package ships -- the source-compatible boundary of SURVEY 8(b).  The reference conv stack runs under cuDNN (TF32 off), the
three native ops run through libflowtrack_b200.so; results are compared with the CPU oracle on the same weights/inputs."""
import os

import pytest
import torch

from oracle import flow_oracle, ref_loader
from tests import gpu_util
from tests.golden import recipes

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_loader.available(), reason="reference not staged under baseline/_ref")]


def _ref_net(name):
    gpu_util.strict_fp32()
    ref = ref_loader.flow_models("b200")
    orc = recipes.flow_oracle_model(name)
    net = getattr(ref, name)(ref_loader.Args()).eval()
    net.load_state_dict(orc.state_dict(), strict=True)
    return ref, orc, net.cuda()


def test_overlay_replaces_only_the_ext_modules():
    ref = ref_loader.flow_models("b200")
    import sys
    over = ref_loader.flow_overlay("b200")
    staged = os.path.join(ref_loader.STAGED, "flownet")
    changed = []
    for dirpath, _d, files in os.walk(over):
        for f in files:
            a, b = os.path.join(dirpath, f), os.path.join(staged, os.path.relpath(os.path.join(dirpath, f), over))
            if not os.path.exists(b) or open(a, "rb").read() != open(b, "rb").read():
                changed.append(os.path.relpath(a, over))
    assert sorted(changed) == sorted("networks/%s_package/_ext/%s/__init__.py" % (op, op) for op in ("correlation", "resample2d", "channelnorm"))
    assert ref.__file__.startswith(over)


@pytest.mark.parametrize("name,hw", [("FlowNet2C", (64, 128)), ("FlowNet2C", (384, 512)), ("FlowNet2CS", (64, 128)), ("FlowNet2CSS", (128, 192))])
def test_reference_modules_over_b200_shims_match_oracle(name, hw):
    """FlowNet2C at 384x512 reaches the tensor-core correlation through the reference's Correlation_forward_cuda signature;
    FlowNet2CS/CSS add the reference's Resample2dFunction / ChannelNormFunction over ft_resample2d_forward / ft_channelnorm_forward."""
    ref, orc, net = _ref_net(name)
    x = recipes.flow_input(1, hw, seed=31)
    with torch.no_grad():
        want = orc(x)
        got = net(x.cuda()).cpu()
    err = (got - want).abs().max().item()
    print("%s %s over the b200 _ext shims: max-abs %.3g on flows up to %.1f px" % (name, hw, err, want.abs().max().item()))
    assert got.shape == want.shape
    assert err <= gpu_util.flow_tol(want), err


def test_reference_functions_raise_like_the_reference_on_cpu_tensors():
    ref = ref_loader.flow_models("b200")
    import importlib
    import sys
    over = ref_loader.flow_overlay("b200")
    sys.path[:0] = [over, os.path.join(over, "networks")]
    try:
        mod = importlib.import_module("networks.resample2d_package.modules.resample2d")
        with pytest.raises(RuntimeError):
            mod.Resample2d()(torch.zeros(1, 3, 8, 8), torch.zeros(1, 2, 8, 8))       # WITH_CUDA-only op: no CPU fallback
    finally:
        del sys.path[:2]
        ref_loader._purge()
