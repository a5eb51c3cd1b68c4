"""Drop-in for the reference's lib/flownet/networks/correlation_package/_ext/correlation/__init__.py (:2-12).

The reference file wraps the cffi library built by build.py (torch.utils.ffi, gone from PyTorch) and exports the C-ABI
names of src/correlation_cuda.h:1-18.  This file exports the same names, backed by libflowtrack_b200.so through ctypes, so
the reference's own, unmodified functions/correlation.py (:28-36) and modules/correlation.py run on top of it.
Copy it over the reference file; nothing else in the reference changes.  Forward only (inference path)."""
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


_so.ft_correlation_shape.argtypes = [_i] * 7 + [ctypes.POINTER(_i)] * 3
_so.ft_correlation_workspace_bytes.restype, _so.ft_correlation_workspace_bytes.argtypes = ctypes.c_size_t, [_i] * 4
_so.ft_correlation_forward_ws.argtypes = [_vp, _vp, _vp] + [_i] * 11 + [_vp, ctypes.c_size_t, _vp]

__all__ = ["Correlation_forward_cuda", "Correlation_backward_cuda"]


def Correlation_forward_cuda(input1, input2, rbot1, rbot2, output, pad_size, kernel_size, max_displacement, stride1, stride2, corr_type_multiply):
    """src/correlation_cuda.c:11-93.  As there, the callee sizes `output` (:36-38) and the scratch tensors: rbot1 becomes
    the workspace of the tensor-core kernel (the reference keeps its zero-padded NHWC copies in rbot1/rbot2, :39-42),
    rbot2 stays empty.  Returns 1 on success like the reference (:90)."""
    _cuda_f32(input1, input2)
    B, C, H, W = input1.shape
    with torch.cuda.device(input1.device):
        oc, oh, ow = _i(), _i(), _i()
        _check(_so.ft_correlation_shape(H, W, pad_size, kernel_size, max_displacement, stride1, stride2, oc, oh, ow))
        output.resize_(B, oc.value, oh.value, ow.value)
        need = _so.ft_correlation_workspace_bytes(B, C, H, W)
        rbot1.resize_((need + 3) // 4)                  # fp32 elements; fresh CUDA allocations are >= 256-byte aligned
        _check(_so.ft_correlation_forward_ws(input1.data_ptr(), input2.data_ptr(), output.data_ptr(), B, C, H, W,
                                             pad_size, kernel_size, max_displacement, stride1, stride2,
                                             0, 0,             # no fused LeakyReLU (FlowNetC applies its own, FlowNetC.py:87); FT_PREC_SPLIT
                                             rbot1.data_ptr(), rbot1.numel() * 4, _stream()))
    return 1


def Correlation_backward_cuda(*args):
    raise NotImplementedError("flowtrack_b200 implements the inference path only (no correlation backward)")
