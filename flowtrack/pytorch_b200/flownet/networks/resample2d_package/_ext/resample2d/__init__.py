"""Drop-in for the reference's lib/flownet/networks/resample2d_package/_ext/resample2d/__init__.py (:2-12): exports the
C-ABI names of src/Resample2d_cuda.h:1-3 backed by libflowtrack_b200.so, so the reference's unmodified
functions/resample2d.py (:8-21) and modules/resample2d.py run on top of it.  Forward only."""
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


_so.ft_resample2d_forward.argtypes = [_vp, _vp, _vp] + [_i] * 4 + [_vp]

__all__ = ["Resample2d_cuda_forward", "Resample2d_cuda_backward"]


def Resample2d_cuda_forward(input1, input2, output, kernel_size):
    """src/Resample2d_cuda.c:8-11 -> Resample2d_kernel.cu:20-66.  `output` [B,C,H,W] is allocated by the caller
    (functions/resample2d.py:17); kernel_size must be 1 (all the reference ever passes)."""
    if kernel_size != 1:
        raise RuntimeError("Resample2d: only kernel_size=1 is implemented")
    _cuda_f32(input1, input2, output)
    B, C, H, W = output.shape
    if tuple(input1.shape) != (B, C, H, W) or tuple(input2.shape) != (B, 2, H, W):
        raise RuntimeError("Resample2d: expected input1 %s and a [B,2,H,W] flow, got %s / %s" % ((B, C, H, W), tuple(input1.shape), tuple(input2.shape)))
    with torch.cuda.device(input1.device):
        _check(_so.ft_resample2d_forward(input1.data_ptr(), input2.data_ptr(), output.data_ptr(), B, C, H, W, _stream()))
    return 1


def Resample2d_cuda_backward(*args):
    raise NotImplementedError("flowtrack_b200 implements the inference path only (no resample backward)")
