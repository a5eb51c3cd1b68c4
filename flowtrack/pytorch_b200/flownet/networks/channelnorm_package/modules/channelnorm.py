"""ChannelNorm module (channelnorm_package/modules/channelnorm.py:8-13): L2 norm over the channel dim -> [B,1,H,W];
computed by ft_channelnorm_forward.  As in the reference kernel, norm_deg is accepted and ignored (always L2)."""
import torch
from torch.nn.modules.module import Module

from ..... import _lib, engine


class ChannelNorm(Module):
    def __init__(self, norm_deg=2):
        super().__init__()
        self.norm_deg = norm_deg

    @torch.no_grad()
    def forward(self, input1):
        x = input1.contiguous()
        if not x.is_cuda or x.dtype != torch.float32:
            raise RuntimeError("ChannelNorm needs a CUDA float32 input (no CPU path)")
        b, c, h, w = x.shape
        with torch.cuda.device(x.device):
            out = torch.empty(b, 1, h, w, device=x.device, dtype=torch.float32)
            _lib.check(_lib.lib().ft_channelnorm_forward(x.data_ptr(), out.data_ptr(), b, c, h, w, engine.stream_ptr()), "channelnorm")
        return out
