"""Drop-in for the reference's lib/flownet/networks/channelnorm_package/_ext/channelnorm/__init__.py (:2-12): exports the
C-ABI names of src/ChannelNorm_cuda.h:1-3 backed by libflowtrack_b200.so, so the reference's unmodified
functions/channelnorm.py (:8-17) and modules/channelnorm.py run on top of it.  Forward only."""
import ctypes
import os

import torch


def _load():
    """libflowtrack_b200.so: $FLOWTRACK_B200_LIB, else the copy inside an importable flowtrack.pytorch_b200 package."""
    path = os.environ.get("FLOWTRACK_B200_LIB")
    if not path:
        from flowtrack.pytorch_b200 import _lib as _ft
        path = _ft.LIB_PATH
    if not os.path.exists(path):
        raise RuntimeError("libflowtrack_b200.so not found (%s): build it with `python -m flowtrack.pytorch_b200.build` "
                           "or point FLOWTRACK_B200_LIB at it; there is no CPU fallback" % path)
    so = ctypes.CDLL(path)
    so.ft_last_error.restype = ctypes.c_char_p
    return so


_so = _load()
_vp, _i = ctypes.c_void_p, ctypes.c_int


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check(rc):
    if rc:                                  # reference: THError("aborting") / THCudaCheck -> Python RuntimeError
        raise RuntimeError(_so.ft_last_error().decode())


def _cuda_f32(*tensors):
    for t in tensors:
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise RuntimeError("expected contiguous CUDA float32 tensors (the reference builds these ops WITH_CUDA only)")


_so.ft_channelnorm_forward.argtypes = [_vp, _vp] + [_i] * 4 + [_vp]

__all__ = ["ChannelNorm_cuda_forward", "ChannelNorm_cuda_backward"]


def ChannelNorm_cuda_forward(input1, output, norm_deg):
    """src/ChannelNorm_cuda.c:8-11 -> ChannelNorm_kernel.cu:19-51 (always the L2 norm: the kernel ignores norm_deg).
    `output` [B,1,H,W] is allocated by the caller (functions/channelnorm.py:11)."""
    _cuda_f32(input1, output)
    B, C, H, W = input1.shape
    if tuple(output.shape) != (B, 1, H, W):
        raise RuntimeError("ChannelNorm: expected output %s, got %s" % ((B, 1, H, W), tuple(output.shape)))
    with torch.cuda.device(input1.device):
        _check(_so.ft_channelnorm_forward(input1.data_ptr(), output.data_ptr(), B, C, H, W, _stream()))
    return 1


def ChannelNorm_cuda_backward(*args):
    raise NotImplementedError("flowtrack_b200 implements the inference path only (no channelnorm backward)")
