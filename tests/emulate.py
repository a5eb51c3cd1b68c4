"""Test-side CPU emulation of the packed conv layers: executes exactly what ft_conv2d_fprop is given
(folded scale/shift, tap-ordered GEMM weights, the 4-phase deconv decomposition, the tile image) with
plain torch ops, so the host-side transforms are validated without a GPU (BASELINE config 1)."""
import numpy as np
import torch

from flowtrack.pytorch_b200 import _lib


def tap_table(layer):
    """Mirror of the tap tables built in ft_conv2d_fprop (csrc/ft_conv.cu)."""
    if layer.kind == _lib.KIND_CONV:
        return [([(r - layer.pad, s - layer.pad) for r in range(layer.kh) for s in range(layer.kw)], (0, 0))], layer.stride, 1
    phases = []
    for py in range(2):
        for px in range(2):
            phases.append(([(py - a, px - b) for a in range(2) for b in range(2)], (py, px)))
    return phases, 1, 2


def emulate_layer(layer, x, res=None):
    """x: [N,H,W,cin_p] fp32 NHWC -> [N,Ho,Wo,cout] fp32 NHWC using layer.host (needs keep_host=True)."""
    wg, sc, sh = (torch.from_numpy(layer.host[k]).double() for k in ("w_gemm", "scale", "shift"))
    N, H, W, C = x.shape
    assert C == layer.cin_p
    phases, in_stride, out_mul = tap_table(layer)
    Ho, Wo = layer.out_hw(H, W)
    Qh, Qw = Ho // out_mul, Wo // out_mul
    out = torch.zeros(N, Ho, Wo, layer.cout_pad, dtype=torch.float64)
    P = 64
    xp = torch.nn.functional.pad(x.double(), (0, 0, P, P, P, P))
    for ph, (taps, (oy, ox)) in enumerate(phases):
        cols = []
        for t, (dy, dx) in enumerate(taps):
            ys = torch.arange(Qh) * in_stride + dy + P
            xs = torch.arange(Qw) * in_stride + dx + P
            cols.append(torch.nn.functional.pad(xp[:, ys][:, :, xs], (0, layer.cin_k - C)))   # per-tap channel padding of the K layout
            if layer.k_mode == _lib.K_ROW and (t + 1) % layer.kw == 0:                          # FT_K_ROW: every kernel row is padded to 64
                cols.append(torch.zeros(N, Qh, Qw, layer.row_pad, dtype=torch.float64))
        A = torch.cat(cols, dim=-1)
        out[:, oy::out_mul, ox::out_mul] = A @ wg[ph][:A.shape[-1]]
    out = out * sc + sh
    out = out[..., :layer.cout]
    if res is not None:
        out = out + res.double()
    if layer.act == _lib.ACT_RELU:
        out = out.clamp_min(0)
    elif layer.act == _lib.ACT_LEAKY:
        out = torch.where(out > 0, out, out * layer.slope)
    return out.float()


def untile(layer, with_residual_blocks=False):
    """Inverse of engine.tile_weights: uint8 image -> fp32 [P, k_pad, cout_pad] (hi + lo).  With
    `with_residual_blocks` also returns the appended residual k-blocks as [P, nt, res_blocks*64, block_n]."""
    bn = layer.block_n
    P = layer.host["w_gemm"].shape[0]
    kb = layer.k_pad // 64 + layer.res_blocks
    t = layer.w_tile_host.view(np.float16).reshape(P, layer.cout_pad // bn, kb, 2, bn, 8, 8)
    u = np.empty_like(t)
    r = np.arange(bn)
    for j in range(8):
        u[:, :, :, :, r, j, :] = t[:, :, :, :, r, j ^ (r % 8), :]
    u = u.astype(np.float32)
    u = u[:, :, :, 0] + u[:, :, :, 1]                  # [P, nt, kb, r, j, e]
    main = u[:, :, :layer.k_pad // 64].transpose(0, 2, 4, 5, 1, 3).reshape(P, layer.k_pad, layer.cout_pad)
    if not with_residual_blocks:
        return main
    resb = u[:, :, layer.k_pad // 64:].transpose(0, 1, 2, 4, 5, 3).reshape(P, layer.cout_pad // bn, layer.res_blocks * 64, bn)
    return main, resb


def nhwc(x, cpad):
    x = x.permute(0, 2, 3, 1)
    return torch.nn.functional.pad(x, (0, cpad - x.shape[-1])).contiguous()


def emulate_pose(model, x):
    """Walk DeconvResnet exactly like DeconvResnet._build_plan does, on the CPU."""
    L = model._pack(None, keep_host=True)
    t = emulate_layer(L["stem"], nhwc(x, 8))
    t = torch.nn.functional.max_pool2d(t.permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
    for blk in L["blocks"]:
        a = emulate_layer(blk["c2"], emulate_layer(blk["c1"], t))
        r = t if blk["down"] is None else emulate_layer(blk["down"], t)
        t = emulate_layer(blk["c3"], a, res=r)
    for dl in L["deconv"]:
        t = emulate_layer(dl, t)
    return emulate_layer(L["head"], t).permute(0, 3, 1, 2).contiguous()
