"""oracle/ref_loader.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Imports the UNMODIFIED reference Python staged under baseline/_ref/ by tools/stage_reference.py (the GPU box has no
/root/reference; the staged copy travels with the snapshot).  Used by tests/, by bench.py's reference arm and by its
`reference_gpu` leg -- never by the product path.

  pose_models()            lib/pose/models            (pose_deconv.py:165-179 `deconv` factory; imports cleanly)
  flow_models(ext)         lib/flownet/model/models   through an overlay of the staged tree in which ONLY the three generated
                           `_ext/<op>/__init__.py` files (torch.utils.ffi wrappers of a cffi build that no longer exists) are
                           replaced:  ext="b200"   -> the drop-in shims shipped in flowtrack/pytorch_b200/flownet/networks/*/_ext
                                                      (libflowtrack_b200.so underneath; needs a GPU to call)
                                      ext="oracle" -> CPU stubs over oracle/ops_oracle.py (the C restatement)
                           so the reference's own modules/*.py and functions/*.py run unmodified on top of either.
"""
import importlib
import os
import shutil
import sys
import tempfile
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, "baseline", "_ref", "lib")
_TOP_LEVEL = ("models", "utils", "model", "networks", "correlation_package", "submodules", "FlowNetC", "FlowNetS", "FlowNetSD", "FlowNetFusion")
_OVERLAYS = {}

_ORACLE_STUBS = {
    "correlation": '''from oracle import ops_oracle
def Correlation_forward_cuda(in1, in2, r1, r2, out, pad, k, md, s1, s2, mult):
    res = ops_oracle.correlation_c(in1.cpu(), in2.cpu(), pad, k, md, s1, s2)
    out.resize_(res.shape).copy_(res)
    return 1
Correlation_backward_cuda = None
''',
    "resample2d": '''from oracle import ops_oracle
def Resample2d_cuda_forward(in1, in2, out, k):
    out.copy_(ops_oracle.resample2d_c(in1.cpu(), in2.cpu()))
    return 1
Resample2d_cuda_backward = None
''',
    "channelnorm": '''from oracle import ops_oracle
def ChannelNorm_cuda_forward(in1, out, deg):
    out.copy_(ops_oracle.channelnorm_c(in1.cpu()))
    return 1
ChannelNorm_cuda_backward = None
''',
}


def available():
    return os.path.isdir(os.path.join(STAGED, "pose", "models")) and os.path.isdir(os.path.join(STAGED, "flownet", "model"))


def _purge():
    for k in [k for k in sys.modules if k.split(".")[0] in _TOP_LEVEL]:
        del sys.modules[k]


def _import_with(paths, name):
    """Import top-level module `name` with `paths` in front of sys.path; the reference's top-level names (models, utils,
    networks, ...) are purged before and after so pose and flow (which both own `utils`) never see each other."""
    _purge()
    saved = list(sys.path)
    sys.path[:0] = paths
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            mod = importlib.import_module(name)
    finally:
        sys.path[:] = saved
        _purge()
    return mod


def pose_models():
    if not available():
        raise RuntimeError("the reference is not staged under baseline/_ref (run tools/stage_reference.py in the build container)")
    return _import_with([os.path.join(STAGED, "pose")], "models")


def flow_overlay(ext):
    """A scratch copy of the staged lib/flownet whose three `_ext/<op>/__init__.py` are replaced (see module docstring)."""
    if ext in _OVERLAYS:
        return _OVERLAYS[ext]
    if not available():
        raise RuntimeError("the reference is not staged under baseline/_ref (run tools/stage_reference.py in the build container)")
    tmp = tempfile.mkdtemp(prefix="ft_ref_flownet_%s_" % ext)
    dst = os.path.join(tmp, "flownet")
    shutil.copytree(os.path.join(STAGED, "flownet"), dst)
    for op in ("correlation", "resample2d", "channelnorm"):
        target = os.path.join(dst, "networks", "%s_package" % op, "_ext", op, "__init__.py")
        os.makedirs(os.path.dirname(target), exist_ok=True)
        if ext == "b200":
            shutil.copyfile(os.path.join(ROOT, "flowtrack", "pytorch_b200", "flownet", "networks", "%s_package" % op, "_ext", op, "__init__.py"), target)
        elif ext == "oracle":
            with open(target, "w") as f:
                f.write(_ORACLE_STUBS[op])
        else:
            raise ValueError(ext)
    _OVERLAYS[ext] = dst
    return dst


def flow_models(ext="oracle"):
    """The reference's lib/flownet/model/models.py (FlowNet2, FlowNet2S, FlowNet2C, ...), sys.path laid out as its own
    tools do (tools/flownet/_init_paths.py: lib/flownet and lib/flownet/networks)."""
    dst = flow_overlay(ext)
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    return _import_with([dst, os.path.join(dst, "networks")], "model.models")


class Args:
    """The attributes the reference constructors read (models.py:21-26): rgb_max, fp16, grads."""
    rgb_max, fp16 = 255.0, False

    def __init__(self):
        self.grads = {}
