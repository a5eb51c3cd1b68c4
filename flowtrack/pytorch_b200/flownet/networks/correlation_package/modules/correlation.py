"""Correlation module: reference constructor / forward signature (correlation_package/modules/correlation.py:8-27),
computed by ft_correlation_forward_ws (tcgen05 banded GEMM for the FlowNetC geometry; the scratch tensor plays the
role of the reference's rbot1/rbot2, functions/correlation.py:28-29) or ft_correlation_forward (CUDA-core kernel for
every other geometry).  Inputs must be contiguous CUDA fp32 NCHW (the reference asserts the same,
functions/correlation.py:17-18).  Forward only: the backward kernels are training code, out of scope."""
import ctypes

import torch
from torch.nn.modules.module import Module

from ..... import _lib, engine


class Correlation(Module):
    def __init__(self, pad_size=0, kernel_size=0, max_displacement=0, stride1=1, stride2=2, corr_multiply=1):
        super().__init__()
        self.pad_size, self.kernel_size, self.max_displacement = pad_size, kernel_size, max_displacement
        self.stride1, self.stride2, self.corr_multiply = stride1, stride2, corr_multiply
        self.precision = _lib.PREC_SPLIT
        self._ws = None

    @torch.no_grad()
    def forward(self, input1, input2, fuse_leaky=False):
        assert input1.is_contiguous() and input2.is_contiguous()
        if not (input1.is_cuda and input2.is_cuda) or input1.dtype != torch.float32 or input2.dtype != torch.float32:
            raise RuntimeError("Correlation needs contiguous CUDA float32 inputs (no CPU path)")
        assert input1.shape == input2.shape
        B, C, H, W = input1.shape
        with torch.cuda.device(input1.device):
            oc, oh, ow = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
            L = _lib.lib()
            _lib.check(L.ft_correlation_shape(H, W, self.pad_size, self.kernel_size, self.max_displacement, self.stride1, self.stride2,
                                              ctypes.byref(oc), ctypes.byref(oh), ctypes.byref(ow)), "correlation_shape")
            out = torch.empty(B, oc.value, oh.value, ow.value, device=input1.device, dtype=torch.float32)
            geom = (self.pad_size, self.kernel_size, self.max_displacement, self.stride1, self.stride2)
            if L.ft_correlation_tc_supported(C, H, W, *geom):
                need = L.ft_correlation_workspace_bytes(B, C, H, W)
                if self._ws is None or self._ws.numel() < need or self._ws.device != input1.device:
                    self._ws = torch.empty(need, dtype=torch.uint8, device=input1.device)
                _lib.check(L.ft_correlation_forward_ws(input1.data_ptr(), input2.data_ptr(), out.data_ptr(), B, C, H, W, *geom, int(fuse_leaky),
                                                       self.precision, self._ws.data_ptr(), self._ws.numel(), engine.stream_ptr()), "correlation")
            else:
                _lib.check(L.ft_correlation_forward(input1.data_ptr(), input2.data_ptr(), out.data_ptr(), B, C, H, W, *geom, int(fuse_leaky),
                                                    engine.stream_ptr()), "correlation")
        return out
