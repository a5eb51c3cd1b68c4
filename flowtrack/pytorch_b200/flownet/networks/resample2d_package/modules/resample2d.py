"""Resample2d module (resample2d_package/modules/resample2d.py:8-14): bilinear backward warp of input1 by the flow
input2, clamp-to-edge; computed by ft_resample2d_forward.  kernel_size must be 1 (all the reference ever uses)."""
import torch
from torch.nn.modules.module import Module

from ..... import _lib, engine


class Resample2d(Module):
    def __init__(self, kernel_size=1):
        super().__init__()
        if kernel_size != 1:
            raise ValueError("Resample2d: only kernel_size=1 is on the path")
        self.kernel_size = kernel_size

    @torch.no_grad()
    def forward(self, input1, input2):
        input1_c = input1.contiguous()
        assert input2.is_contiguous()
        if not (input1_c.is_cuda and input2.is_cuda) or input1_c.dtype != torch.float32 or input2.dtype != torch.float32:
            raise RuntimeError("Resample2d needs CUDA float32 inputs (no CPU path)")
        _, d, _, _ = input1_c.size()
        b, two, h, w = input2.size()
        assert two == 2 and input1_c.shape[0] == b and input1_c.shape[2:] == (h, w)
        with torch.cuda.device(input1_c.device):
            out = torch.empty(b, d, h, w, device=input1.device, dtype=torch.float32)
            _lib.check(_lib.lib().ft_resample2d_forward(input1_c.data_ptr(), input2.data_ptr(), out.data_ptr(), b, d, h, w, engine.stream_ptr()), "resample2d")
        return out
