"""oracle/ref_gpu.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes access to the reference's OWN CUDA kernels compiled unmodified into oracle/_ref/ (oracle/Makefile, target `ref`):
  libref_warp.so          Resample2d_kernel_forward (resample2d_package/src/Resample2d_kernel.cu:188-206 -> kernel :20-66)
                          ChannelNorm_kernel_forward (channelnorm_package/src/ChannelNorm_kernel.cu:82-98 -> kernel :19-51)
  libref_correlation.so   Correlation_forward_cuda_kernel (correlation_package/src/correlation_cuda_kernel.cu:296-369)
Used by the -m gpu tests (bit-level cross-check of the B200 kernels) and by bench.py's `reference_gpu` leg (the on-box
bar of the reference kernels).  Only tests/ and bench.py may import this module."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))


class THCudaTensor(ctypes.Structure):        # mirrors oracle/ref_shims/THC/THC.h
    _fields_ = [("size", ctypes.c_long * 4), ("stride", ctypes.c_long * 4), ("data", ctypes.c_void_p)]


class THCState(ctypes.Structure):
    _fields_ = [("stream", ctypes.c_void_p)]


def _load(name):
    path = os.path.join(HERE, "_ref", name)
    return ctypes.CDLL(path) if os.path.exists(path) else None


def _view(t):
    assert t.is_cuda and t.dim() == 4
    return THCudaTensor((ctypes.c_long * 4)(*t.shape), (ctypes.c_long * 4)(*t.stride()), t.data_ptr())


def warp_lib():
    return _load("libref_warp.so")


def _state():
    import torch
    return THCState(torch.cuda.current_stream().cuda_stream)


def resample2d(img, flow):
    """The reference kernel_Resample2d_updateOutput on contiguous CUDA fp32 [B,C,H,W] / [B,2,H,W]."""
    import torch
    lib = warp_lib()
    out = torch.zeros(flow.shape[0], img.shape[1], flow.shape[2], flow.shape[3], device=img.device)
    st, a, b, o = _state(), _view(img), _view(flow), _view(out)
    lib.Resample2d_kernel_forward(ctypes.byref(st), ctypes.byref(a), ctypes.byref(b), ctypes.byref(o), 1)
    return out


def channelnorm(x):
    import torch
    lib = warp_lib()
    out = torch.zeros(x.shape[0], 1, x.shape[2], x.shape[3], device=x.device)
    st, a, o = _state(), _view(x), _view(out)
    lib.ChannelNorm_kernel_forward(ctypes.byref(st), ctypes.byref(a), ctypes.byref(o), 2)
    return out


def correlation_lib():
    return _load("libref_correlation.so")


def correlation(a, b, out=None, scratch=None, pad=20, ksize=1, max_disp=20, s1=1, s2=2):
    """The reference Correlation_forward_cuda_kernel (FlowNetC geometry by default); `scratch` = (rInput1, rInput2) zero-filled
    NHWC padded copies as Correlation_forward_cuda allocates them (correlation_cuda.c:36-42)."""
    import torch
    lib = correlation_lib()
    B, C, H, W = a.shape
    d = max_disp // s2
    oc = (2 * d + 1) ** 2
    if out is None:
        out = torch.zeros(B, oc, H, W, device=a.device)
    if scratch is None:
        scratch = (torch.zeros(B, H + 2 * pad, W + 2 * pad, C, device=a.device), torch.zeros(B, H + 2 * pad, W + 2 * pad, C, device=a.device))
    vp, ci = ctypes.c_void_p, ctypes.c_int
    lib.Correlation_forward_cuda_kernel.argtypes = [vp] + [ci] * 8 + [vp] + [ci] * 7 + [vp] + [ci] * 5 + [vp, vp] + [ci] * 6 + [vp]
    ok = lib.Correlation_forward_cuda_kernel(out.data_ptr(), B, oc, H, W, *out.stride(), a.data_ptr(), C, H, W, *a.stride(),
                                             b.data_ptr(), C, *b.stride(), scratch[0].data_ptr(), scratch[1].data_ptr(), pad, ksize, max_disp, s1, s2, 1,
                                             ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert ok == 1
    return out
