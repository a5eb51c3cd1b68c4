"""Host side of the conv stack: weight pre-transformation, split-fp16 NHWC activation buffers, and
per-layer dispatch through the C ABI (include/flowtrack_b200.h).

Nothing here computes network arithmetic on the host: the transforms are one-off layout/folding steps
at weight-load time (eval-mode BatchNorm folded into a per-channel scale/shift; the
k4 s2 p1 transposed convolution split into its four output-parity 2x2 sub-convolutions, SURVEY.md
7.3 item 5), and every forward step is a kernel launch in libflowtrack_b200.so.
"""
import ctypes
import hashlib
import json
import os

import numpy as np
import torch

from . import _lib
from ._lib import FtAct, FtConvDesc

BK = 64  # k-block (fp16 elements) = one 128-byte swizzle row
AUTOTUNE_MARGIN = 0.015  # Plan.autotune(): a candidate tiling replaces the static choice only if it is more than this fraction faster


def round_up(a, b):
    return (a + b - 1) // b * b


def pad8(c):
    return round_up(c, 8)


# --------------------------------------------------------------------------------------------------
# weight transforms (host, numpy) -- exercised on CPU by tests/test_engine_cpu.py
# --------------------------------------------------------------------------------------------------
def fold_bn(bn):
    """Eval-mode BatchNorm2d -> (scale, shift): y = x*scale + shift."""
    s = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)
    t = bn.bias.detach().double() - bn.running_mean.detach().double() * s
    return s.float().cpu().numpy(), t.float().cpu().numpy()


def gemm_weights_conv(weight, cin_k):
    """Conv2d weight [Cout,Cin,kh,kw] -> [1, k_pad, Cout] with k = (r*kw+s)*cin_k + ci  (cin_k >= Cin: the per-tap
    channel count of the K layout -- the padded view width for FT_K_PACKED, that rounded up to 64 for FT_K_TAP64)."""
    w = weight.detach().float().cpu().numpy()
    cout, cin, kh, kw = w.shape
    k = kh * kw * cin_k
    g = np.zeros((1, round_up(k, BK), cout), np.float32)
    g[0, :k].reshape(kh * kw, cin_k, cout)[:, :cin] = w.transpose(2, 3, 1, 0).reshape(kh * kw, cin, cout)
    return g


def gemm_weights_conv_rows(weight, cin_p):
    """FT_K_ROW layout: Conv2d weight [Cout,Cin,kh,kw] -> [1, kh*64, Cout] with k = r*64 + s*cin_p + ci (kw*cin_p <= 64)."""
    w = weight.detach().float().cpu().numpy()
    cout, cin, kh, kw = w.shape
    assert kw * cin_p <= BK
    g = np.zeros((1, kh, BK, cout), np.float32)
    g[0, :, :kw * cin_p].reshape(kh, kw, cin_p, cout)[:, :, :cin] = w.transpose(2, 3, 1, 0)
    return g.reshape(1, kh * BK, cout)


def gemm_weights_deconv(weight, cin_k):
    """ConvTranspose2d(k4,s2,p1) weight [Cin,Cout,4,4] -> [4, k_pad, Cout].
    out[oy] += in[iy]*w[ky] with oy = 2*iy - 1 + ky, so output parity py only sees ky = 1 - py + 2a (a=0,1), read at
    iy = qy + py - a.  phase = py*2+px, k = (a*2+b)*cin_k + ci  (matches the tap table in ft_conv2d_fprop)."""
    w = weight.detach().float().cpu().numpy()
    cin, cout = w.shape[:2]
    assert w.shape[2:] == (4, 4)
    k = 4 * cin_k
    g = np.zeros((4, round_up(k, BK), cout), np.float32)
    for py in range(2):
        for px in range(2):
            v = g[py * 2 + px, :k].reshape(4, cin_k, cout)
            for a in range(2):
                for b in range(2):
                    v[a * 2 + b, :cin] = w[:, :, 1 - py + 2 * a, 1 - px + 2 * b]
    return g


def normalise_columns(g, scale, max_exp=None):
    """Scale every output channel by a power of two f so max|w| lands in [1024, 2048): keeps the fp16 `lo`
    halves out of the subnormal range.  Exact (powers of two), undone through `scale`.  Returns (g*f, scale/f, f).
    `max_exp` caps log2 f (residual layers put f itself into fp16 identity blocks)."""
    m = np.abs(g).max(axis=(0, 1))
    e = np.where(m > 0, 10 - np.floor(np.log2(np.maximum(m, 1e-38))), 0.0)
    if max_exp is not None:
        e = np.minimum(e, max_exp)
    f = np.exp2(e).astype(np.float32)
    return g * f[None, None, :], (scale / f).astype(np.float32), f


def box_tiles(qw, qh, n):
    """Number of 128-pixel box tiles ft_conv2d_fprop cuts an n x qh x qw output grid into (mirror of choose_box in ft_conv.cu)."""
    best = None
    for lw in range(7, -1, -1):
        for lh in range(7 - lw, -1, -1):
            w, h = 1 << lw, 1 << lh
            t = -(-qw // w) * -(-qh // h) * -(-n // (128 // (w * h)))
            if best is None or t < best:
                best = t
    return best


def choose_block_n(cout, precision=_lib.PREC_SPLIT):
    import os
    force = int(os.environ.get("FT_DEV_BLOCK_N", "0"))      # dev knob for tile-shape experiments (tools/dev/layer_times.py)
    if force and cout >= force:
        return force
    if cout >= 128:
        return 128
    if cout > 32:
        return 64
    if cout > 16:
        return 32
    return 16


def tile_weights(g, block_n, res_f=None):
    """[P, k_pad, cout_pad] fp32 -> uint8 image: for each (phase, n_tile, k_block) a [2 (hi,lo)][block_n][64]
    fp16 block in the tcgen05 K-major SWIZZLE_128B layout: 16-byte chunk j of row r sits at chunk j ^ (r % 8).
    `res_f` (per-channel powers of two) appends ceil(block_n/64) residual k-blocks to every (phase, n_tile): block j is
    diag(res_f) restricted to channels [64j, 64j+64) of the tile, so the kernel's  acc += residual_tile @ block^T
    adds residual*f to the accumulator (flowtrack_b200.h, ft_conv_desc.res_blocks)."""
    P, kp, cp = g.shape
    assert kp % BK == 0 and cp % block_n == 0
    if res_f is not None:
        nres, nt = -(-block_n // BK), cp // block_n
        ext = np.zeros((P, nt, kp + nres * BK, block_n), np.float32)       # per n-tile K extension
        ext[:, :, :kp] = g.reshape(P, kp, nt, block_n).transpose(0, 2, 1, 3)
        fr = np.asarray(res_f, np.float32).reshape(nt, block_n)
        assert np.all(np.abs(fr) <= 32768.0)
        for r in range(block_n):
            ext[:, :, kp + r, r] = fr[None, :, r]                          # k = kp + (64j + kk) <-> channel r = 64j + kk
        g = ext.transpose(0, 2, 1, 3).reshape(P, kp + nres * BK, cp)
        kp = kp + nres * BK
    hi = g.astype(np.float16)
    lo = (g - hi.astype(np.float32)).astype(np.float16)
    t = np.stack((hi, lo), 0)                                              # [2, P, k, c]
    t = t.reshape(2, P, kp // BK, 8, 8, cp // block_n, block_n)            # [2, P, kb, j, e, nt, r]
    t = t.transpose(1, 5, 2, 0, 6, 3, 4)                                   # [P, nt, kb, 2, r, j, e]
    out = np.empty_like(t)
    r = np.arange(block_n)
    for j in range(8):
        out[:, :, :, :, r, j ^ (r % 8), :] = t[:, :, :, :, r, j, :]
    return np.ascontiguousarray(out).view(np.uint8).reshape(-1)


# --------------------------------------------------------------------------------------------------
# activations
# --------------------------------------------------------------------------------------------------
class Act:
    """Split-fp16 NHWC activation buffer on the GPU: value = hi + lo.  `view(coff, c)` selects a channel slice."""

    def __init__(self, n, h, w, c, device, split=True, zero=False, xpad=(0, 0)):
        """`xpad` = (left, right) zero columns physically stored in every row (the FT_K_ROW stem convolutions read their
        zero border from them); w stays the image width, struct() reports the physical row pitch."""
        self.n, self.h, self.w, self.c = n, h, w, pad8(c)
        self.ld, self.coff = self.c, 0
        self.xl, self.xr = xpad
        mk = torch.zeros if (zero or self.xl or self.xr) else torch.empty
        self.hi = mk((n, h, w + self.xl + self.xr, self.ld), dtype=torch.float16, device=device)
        self.lo = mk((n, h, w + self.xl + self.xr, self.ld), dtype=torch.float16, device=device) if split else None
        self._base = self

    def view(self, coff, c, n0=0, n=None):
        v = object.__new__(Act)
        v.__dict__.update(self.__dict__)
        v.coff, v.c = self.coff + coff, pad8(c)
        v._n0, v.n = getattr(self, "_n0", 0) + n0, (self.n - n0 if n is None else n)
        assert v.coff % 8 == 0 and v.coff + v.c <= self.ld
        return v

    def struct(self):
        n0 = getattr(self, "_n0", 0)
        wp = self.w + self.xl + self.xr
        off = n0 * self.h * wp * self.ld * 2
        return FtAct(self.hi.data_ptr() + off, (self.lo.data_ptr() + off) if self.lo is not None else None,
                     self.n, self.h, wp, self.c, self.ld, self.coff)

    def to_nchw(self, c=None):
        """Debug/test helper: materialise the view as fp32 NCHW."""
        n0 = getattr(self, "_n0", 0)
        c = self.c if c is None else c
        xs = slice(self.xl, self.xl + self.w)
        t = self.hi[n0:n0 + self.n, :, xs, self.coff:self.coff + c].float()
        if self.lo is not None:
            t = t + self.lo[n0:n0 + self.n, :, xs, self.coff:self.coff + c].float()
        return t.permute(0, 3, 1, 2).contiguous()


NULL_ACT = FtAct(None, None, 0, 0, 0, 0, 0, 0)


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


# --------------------------------------------------------------------------------------------------
# one fused conv layer
# --------------------------------------------------------------------------------------------------
class ConvLayer:
    """conv / deconv(k4,s2,p1) + per-channel scale/shift + optional residual + activation, weights resident on the GPU."""

    def __init__(self, kind, weight, scale, shift, stride=1, pad=0, act=_lib.ACT_NONE, slope=0.1, cin_p=None, device=None, block_n=None, keep_host=False,
                 residual=False, k_mode=None, pad_hi=None, flops_per_pixel=None):
        """pad_hi: zero padding after the last row/column when it differs from `pad` (space-to-depth stems);
        flops_per_pixel: algorithmic FLOPs per output pixel when the packed weights carry structural zeros."""
        if kind == _lib.KIND_CONV:
            self.cout, cin, self.kh, self.kw = weight.shape
        else:
            cin, self.cout, self.kh, self.kw = weight.shape
        self.kind, self.stride, self.pad, self.act, self.slope = kind, stride, pad, act, slope
        self.pad_hi = pad if pad_hi is None else pad_hi
        self.flops_per_pixel = flops_per_pixel
        self.cin, self.cin_p = cin, cin_p or pad8(cin)
        assert self.cin_p % 8 == 0 and self.cin_p >= cin
        # K layout: small-Cin plain convolutions (the stems) pack K across taps and are gathered by cp.async;
        # everything else gives each tap a whole number of 64-channel blocks so a k-block is one TMA box.
        # The 7x7 stems (kw*Cin between 49 and 64) use FT_K_ROW: one k-block per kernel row of an x-padded input.
        packed = kind == _lib.KIND_CONV and not residual and (self.cin_p < 32 or stride not in (1, 2))
        rows = packed and stride in (1, 2) and 48 < self.kw * self.cin_p <= BK and k_mode != _lib.K_PACKED
        self.k_mode = _lib.K_ROW if rows else (_lib.K_PACKED if packed else _lib.K_TAP64)
        self.cin_k = round_up(self.cin_p, BK) if self.k_mode == _lib.K_TAP64 else self.cin_p
        self.row_pad = BK - self.kw * self.cin_p if rows else 0
        # zero columns the input rows must carry for FT_K_ROW (left, right)
        self.in_xpad = (round_up(pad, 4), round_up(pad + (BK - self.kw * self.cin_p + self.cin_p - 1) // self.cin_p, 4)) if rows else (0, 0)
        self.residual = residual
        if rows:
            g = gemm_weights_conv_rows(weight, self.cin_p)
        else:
            g = gemm_weights_conv(weight, self.cin_k) if kind == _lib.KIND_CONV else gemm_weights_deconv(weight, self.cin_k)
        self.block_n = block_n or choose_block_n(self.cout)                    # default tile width (single-CTA tiles)
        # cout is padded so that every tile width the scheduler may pick (64 / 128 / 256) divides it
        self.cout_pad = round_up(self.cout, 256 if (self.cout > 128 and block_n is None) else self.block_n)
        self.k_pad = g.shape[1]
        self.res_blocks = -(-self.block_n // BK) if residual else 0            # of the default tile image
        gp = np.zeros((g.shape[0], self.k_pad, self.cout_pad), np.float32)
        sc = np.zeros(self.cout_pad, np.float32)
        sh = np.zeros(self.cout_pad, np.float32)
        gp[:, :, :self.cout], sc[:self.cout], sh[:self.cout] = g, scale, shift
        if residual:
            # out = (acc + res*f) / f + shift needs 1/scale == f, a power of two: fold BatchNorm's scale into the weights
            gp = (gp.astype(np.float64) * sc.astype(np.float64)[None, None, :]).astype(np.float32)
            sc = np.where(np.arange(self.cout_pad) < self.cout, 1.0, 0.0).astype(np.float32)
        gp, sc, f = normalise_columns(gp, sc, max_exp=14 if residual else None)
        self._res_f = np.where(sc != 0, f, 0.0) if residual else None
        self._gp = gp                                                          # host fp32 GEMM weights: tile images are cut from it on demand
        self._tiles = {}                                                       # block_n -> device tile image
        self.device = device
        self.host = dict(w_gemm=gp, scale=sc, shift=sh) if keep_host else None
        self.w_tile_host = tile_weights(gp, self.block_n, self._res_f) if keep_host else None
        if device is not None:
            self.w_gemm = torch.from_numpy(gp).to(device)
            self.scale, self.shift = torch.from_numpy(sc).to(device), torch.from_numpy(sh).to(device)

    def tile_image(self, block_n):
        """Device tile image of the weights for tiles `block_n` channels wide (built once per width)."""
        if block_n not in self._tiles:
            assert self.cout_pad % block_n == 0
            self._tiles[block_n] = torch.from_numpy(tile_weights(self._gp, block_n, self._res_f)).to(self.device)
        return self._tiles[block_n]

    def choose_tiling(self, n, h, w, precision=_lib.PREC_SPLIT):
        """(block_n, cta_pair) for an n x h x w input.  Measured on B200 with the warp-uniform MMA issue path
        (profiles/r01i_conv_tilings.txt): CTA pairs (tcgen05 cta_group::2, 256 pixels x block_n per pair, each CTA staging half
        of the weight tile) win on every TMA-fed layer -- deconv.6 266 -> 236 us, layer3 3x3 57 -> 46, layer4 3x3 58 -> 45,
        FlowNet conv3_1 155 -> 133 -- because the single-CTA main loop is bound by the shared-memory fill (64 KiB per k-block
        against 768 tensor-pipe cycles).  256-wide tiles lose with the chunked accumulation.  FT_DEV_PAIR=0 turns pairs off."""
        import os
        pair_ok = self.k_mode != _lib.K_PACKED and self.block_n >= 64 and os.environ.get("FT_DEV_PAIR", "1") != "0"
        if pair_ok:
            ho, wo = self.out_hw(h, w)
            qh, qw = (h, w) if self.kind == _lib.KIND_DECONV else (ho, wo)
            if box_tiles(qw, qh, n) >= 2:
                return self.block_n, 1
        return self.block_n, 0

    def out_hw(self, h, w):
        if self.kind == _lib.KIND_DECONV:
            return 2 * h, 2 * w
        return (h + self.pad + self.pad_hi - self.kh) // self.stride + 1, (w + self.pad + self.pad_hi - self.kw) // self.stride + 1

    def flops(self, n, h, w):
        """Algorithmic FLOPs (2 x MACs over the real, un-padded channels) of one forward on an n x h x w input."""
        if self.kind == _lib.KIND_DECONV:
            return 2 * n * h * w * self.cin * self.cout * self.kh * self.kw
        ho, wo = self.out_hw(h, w)
        if self.flops_per_pixel is not None:
            return n * ho * wo * self.flops_per_pixel
        return 2 * n * ho * wo * self.cin * self.cout * self.kh * self.kw

    def tile_counts(self, n, h, w, bn, pair):
        """(work tiles, CTA tiles) ft_conv2d_fprop cuts an n x h x w input into with tile width bn (pair: 256-pixel tiles on CTA pairs)."""
        ho, wo = self.out_hw(h, w)
        qh, qw = (h, w) if self.kind == _lib.KIND_DECONV else (ho, wo)
        mt = box_tiles(qw, qh, n)
        phases, nt = (4 if self.kind == _lib.KIND_DECONV else 1), self.cout_pad // bn
        if pair == 2:                            # clusters of four: a work item = 256 pixels x two n-tiles
            return phases * ((mt + 1) // 2) * (nt // 2), phases * 2 * ((mt + 1) // 2) * nt
        return phases * ((mt + 1) // 2 if pair else mt) * nt, phases * (2 * ((mt + 1) // 2) if pair else mt) * nt

    def desc(self, x, out=None, out_f32=None, res=None, precision=_lib.PREC_SPLIT, impl=_lib.IMPL_TCGEN05, tiling=None, splitk_ws=None):
        """`tiling` = (block_n, cta_pair[, epilogue_buffers[, splitk]]); default: choose_tiling() for this input shape.
        splitk > 1 needs `splitk_ws` = (uint8 workspace tensor, int32 counter tensor) owned by the caller (engine.Plan)."""
        assert x.c == self.cin_p, "input view has %d channels, layer expects %d" % (x.c, self.cin_p)
        assert (res is not None) == self.residual, "layer was packed %s residual blocks" % ("with" if self.residual else "without")
        assert (x.xl, x.xr) == self.in_xpad, "layer expects input rows padded by %s zero columns, view has %s" % (self.in_xpad, (x.xl, x.xr))
        d = FtConvDesc()
        d.in_ = x.struct()
        d.out = out.struct() if out is not None else NULL_ACT
        d.res = res.struct() if res is not None else NULL_ACT
        d.out_f32 = out_f32.data_ptr() if out_f32 is not None else None
        tiling = tiling if tiling is not None else self.choose_tiling(x.n, x.h, x.w, precision)
        bn, pair = tiling[0], tiling[1]
        d.epilogue_buffers = tiling[2] if len(tiling) > 2 else 1
        d.splitk = tiling[3] if len(tiling) > 3 else 0
        d.ctas_per_sm = tiling[4] if len(tiling) > 4 else 1
        if d.splitk > 1:
            ws, cnt = splitk_ws
            d.splitk_ws, d.splitk_ws_bytes = ws.data_ptr(), ws.numel()
            d.splitk_counters, d.splitk_counters_len = cnt.data_ptr(), cnt.numel()
        d.w_tile, d.w_gemm = self.tile_image(bn).data_ptr(), self.w_gemm.data_ptr()
        d.scale, d.shift = self.scale.data_ptr(), self.shift.data_ptr()
        d.kind, d.kh, d.kw, d.stride, d.pad = self.kind, self.kh, self.kw, self.stride, self.pad
        d.pad_hi = self.pad_hi if self.pad_hi != self.pad else -1
        d.cout, d.cout_pad, d.k_pad, d.block_n = self.cout, self.cout_pad, self.k_pad, bn
        d.act, d.slope, d.precision, d.impl = self.act, self.slope, precision, impl
        d.k_mode, d.res_blocks, d.cta_pair = self.k_mode, (-(-bn // BK) if self.residual else 0), pair
        d.in_xpad_l, d.in_xpad_r = x.xl, x.xr
        return d


def conv_from_modules(conv, bn=None, act=_lib.ACT_NONE, cin_p=None, device=None, keep_host=False, residual=False, k_mode=None):
    """nn.Conv2d / nn.ConvTranspose2d (+ eval BatchNorm2d) -> ConvLayer.  `residual`: the layer will be called with a
    residual tensor (bottleneck conv3); its weight image then carries the identity k-blocks the kernel needs."""
    is_deconv = isinstance(conv, torch.nn.ConvTranspose2d)
    cout = conv.out_channels
    if bn is not None:
        scale, shift = fold_bn(bn)
        if conv.bias is not None:
            shift = shift + conv.bias.detach().float().cpu().numpy() * scale
    else:
        scale = np.ones(cout, np.float32)
        shift = conv.bias.detach().float().cpu().numpy() if conv.bias is not None else np.zeros(cout, np.float32)
    if is_deconv:
        assert conv.kernel_size == (4, 4) and conv.stride == (2, 2) and conv.padding == (1, 1) and conv.output_padding == (0, 0)
        assert not residual
        return ConvLayer(_lib.KIND_DECONV, conv.weight, scale, shift, act=act, cin_p=cin_p, device=device, keep_host=keep_host)
    assert conv.kernel_size[0] == conv.kernel_size[1] and conv.stride[0] == conv.stride[1] and conv.padding[0] == conv.padding[1]
    assert conv.dilation == (1, 1) and conv.groups == 1
    return ConvLayer(_lib.KIND_CONV, conv.weight, scale, shift, stride=conv.stride[0], pad=conv.padding[0], act=act,
                     cin_p=cin_p, device=device, keep_host=keep_host, residual=residual, k_mode=k_mode)


def stem_s2d_weights(weight):
    """Conv2d(3, Cout, 7, stride 2, pad 3) weight [Cout,3,7,7] -> the equivalent 4x4 stride-1 (pad 2 before, 1 after) weight
    [Cout,16,4,4] over the space-to-depth image of ft_pack_image_s2d (channel (py*2+px)*4 + c = src[c, 2Y+py, 2X+px]):
    input row 2y-3+ky is s2d row y-2+r with parity py, ky = 2r + py - 1 (ky = -1 does not exist: zero weight)."""
    w = weight.detach().float().cpu()
    cout, cin, kh, kw = w.shape
    assert (cin, kh, kw) == (3, 7, 7)
    w4 = torch.zeros(cout, 16, 4, 4)
    for r in range(4):
        for py in range(2):
            ky = 2 * r + py - 1
            if not 0 <= ky < 7:
                continue
            for s_ in range(4):
                for px in range(2):
                    kx = 2 * s_ + px - 1
                    if 0 <= kx < 7:
                        w4[:, (py * 2 + px) * 4:(py * 2 + px) * 4 + 3, r, s_] = w[:, :, ky, kx]
    return w4


def stem_s2d_from_modules(conv, bn=None, act=_lib.ACT_NONE, device=None, keep_host=False):
    """7x7 stride-2 pad-3 three-channel stem (+ eval BatchNorm) -> the space-to-depth ConvLayer (FT_K_ROW, 4 k-blocks of 64
    instead of 7; input from ft_pack_image_s2d).  Same arithmetic: every product of the original appears exactly once."""
    assert conv.kernel_size == (7, 7) and conv.stride == (2, 2) and conv.padding == (3, 3) and conv.in_channels == 3 and conv.groups == 1
    cout = conv.out_channels
    if bn is not None:
        scale, shift = fold_bn(bn)
        if conv.bias is not None:
            shift = shift + conv.bias.detach().float().cpu().numpy() * scale
    else:
        scale = np.ones(cout, np.float32)
        shift = conv.bias.detach().float().cpu().numpy() if conv.bias is not None else np.zeros(cout, np.float32)
    layer = ConvLayer(_lib.KIND_CONV, stem_s2d_weights(conv.weight), scale, shift, stride=1, pad=2, pad_hi=1, act=act, cin_p=16, device=device,
                      keep_host=keep_host, flops_per_pixel=2 * 3 * 49 * cout)
    assert layer.k_mode == _lib.K_ROW
    return layer


class FlowHead:
    """predict_flow = Conv2d(C, 2, 3, 1, 1) (lib/flownet/networks/submodules.py:20-21) as a 1x1 convolution with 18
    outputs on the tensor cores, part[(r*3+s)*2 + c] = W[c,:,r,s] . x, finished by ft_flow_head's shifted 9-term sum.
    As a 3x3 implicit GEMM the wide concat input would be re-read nine times for two output columns."""

    def __init__(self, conv, device=None, keep_host=False):
        w = conv.weight.detach()
        assert tuple(w.shape[2:]) == (3, 3) and w.shape[0] == 2 and conv.stride == (1, 1) and conv.padding == (1, 1)
        self.cin, self.cout = w.shape[1], 2
        w18 = w.permute(2, 3, 0, 1).reshape(18, self.cin, 1, 1).contiguous()
        self.gemm = ConvLayer(_lib.KIND_CONV, w18, np.ones(18, np.float32), np.zeros(18, np.float32), device=device, keep_host=keep_host)
        self.cin_p = self.gemm.cin_p
        self.bias_host = conv.bias.detach().float().cpu().contiguous() if conv.bias is not None else None
        self.bias = self.bias_host.to(device) if (self.bias_host is not None and device is not None) else None


class FlowUp:
    """upsampled_flowX_to_Y = ConvTranspose2d(2, 2, 4, 2, 1) on the fp32 flow (FlowNetS.py:41-44) -> ft_flow_upsample."""

    def __init__(self, deconv, device=None):
        assert tuple(deconv.weight.shape) == (2, 2, 4, 4) and deconv.stride == (2, 2) and deconv.padding == (1, 1)
        self.w_host = deconv.weight.detach().float().cpu().contiguous()
        self.bias_host = deconv.bias.detach().float().cpu().contiguous() if deconv.bias is not None else None
        self.w = self.w_host.to(device) if device is not None else None
        self.bias = self.bias_host.to(device) if (self.bias_host is not None and device is not None) else None


# --------------------------------------------------------------------------------------------------
# plans: a forward pass = a fixed list of C-ABI calls over pre-allocated buffers (CUDA-graph friendly)
# --------------------------------------------------------------------------------------------------
class Plan:
    def __init__(self, device, precision=_lib.PREC_SPLIT, impl=_lib.IMPL_TCGEN05):
        self.device, self.precision, self.impl = device, precision, impl
        self.calls = []      # (callable, args tuple); args keep ctypes structs alive
        self.keep = []       # buffers referenced by raw pointer
        self.graph = None
        self.taps = {}       # name -> Act / tensor, for layer-by-layer tests
        self.flops = {}      # call index -> algorithmic FLOPs (convolutions only)
        self._tag, self._ntags = 0, 1   # branch (stream) tag of the calls being added; see branch()
        self._tags = {}       # call index -> branch tag (absent = 0, the caller's stream)
        self._streams = {}    # branch tag -> side stream (created lazily)
        self._conv_args = {}  # call index -> (layer, x, out, out_f32, res): what autotune() rebuilds a descriptor from
        self.tuned = {}      # call index -> (block_n, cta_pair, us) chosen by autotune()
        self.alternate_tile_order = os.environ.get("FT_TILE_ORDER", "alternate") == "alternate"
        self._rev = 0        # tile order of the next convolution (alternates: a consumer starts where its producer ended, in L2)

    def act(self, n, h, w, c, zero=False, xpad=(0, 0)):
        a = Act(n, h, w, c, self.device, split=True, zero=zero, xpad=xpad)
        self.keep.append(a)
        return a

    def f32(self, *shape):
        t = torch.empty(shape, dtype=torch.float32, device=self.device)
        self.keep.append(t)
        return t

    def branch(self):
        """Context manager: the calls added inside form a branch that is independent of the calls added after it until the matching
        `join()` -- e.g. a FlowNet decoder level's flow head (1x1 conv on a handful of tiles + two tiny kernels) next to its
        deconvolution, or FlowNetSD next to the FlowNetC -> S -> S chain.  enqueue() issues a branch on a side stream between a fork
        and a join event, so inside a CUDA graph the branches are parallel paths: kernels that cannot fill the GPU overlap.  The
        linear order of `calls` stays a valid schedule (time_calls* and the tests use it).  Branches nest."""
        plan = self

        class _Branch:
            def __enter__(self_b):
                self_b.parent, self_b.tag = plan._tag, plan._ntags
                plan._ntags += 1
                plan.calls.append(("fork", None, (self_b.tag, self_b.parent), None))
                plan._tag = self_b.tag
                return self_b

            def __exit__(self_b, *exc):
                plan._tag = self_b.parent
                return False

            def join(self_b):
                plan.calls.append(("join", None, (self_b.tag, plan._tag), None))
        return _Branch()

    def _add(self, call):
        self.calls.append(call)
        if self._tag:
            self._tags[len(self.calls) - 1] = self._tag

    def conv(self, layer, x, out=None, out_f32=None, res=None, name=None):
        d = layer.desc(x, out, out_f32, res, self.precision, self.impl)
        if self.alternate_tile_order:
            d.tile_order = self._rev
            self._rev ^= 1
        self._add(("conv2d", _lib.lib().ft_conv2d_fprop, (ctypes.byref(d),), d))
        self._conv_args[len(self.calls) - 1] = (layer, x, out, out_f32, res)
        self.flops[len(self.calls) - 1] = layer.flops(x.n, x.h, x.w)
        if name:
            self.taps[name] = out if out is not None else out_f32
        return out if out is not None else out_f32

    def call(self, what, fn, *args):
        self._add((what, fn, args, args))

    def flow_head(self, head, x, out=None, name=None):
        """predict_flow on the split-fp16 view x -> fp32 NCHW flow [B,2,h,w] (a new static tensor unless `out` is given)."""
        part = self.f32(x.n, 18, x.h, x.w)
        self.conv(head.gemm, x, None, out_f32=part)
        flow = out if out is not None else self.f32(x.n, 2, x.h, x.w)
        self.call("flow_head", _lib.lib().ft_flow_head, part.data_ptr(), head.bias.data_ptr() if head.bias is not None else None,
                  flow.data_ptr(), x.n, x.h, x.w)
        self.keep.append(head)
        if name:
            self.taps[name] = flow
        return flow

    def flow_up(self, up, flow, dst):
        """ConvTranspose2d(2,2,4,2,1) of an fp32 NCHW flow into the 8-channel split-fp16 view dst (a concat-buffer slice)."""
        B, _, h, w = flow.shape
        self.call("flow_upsample", _lib.lib().ft_flow_upsample, flow.data_ptr(), up.w.data_ptr(), up.bias.data_ptr() if up.bias is not None else None,
                  B, h, w, dst.struct())
        self.keep.append(up)
        return dst

    def enqueue(self, start=0):
        """Issue every call of the plan (from call index `start`: a caller that fills the plan's packed input itself skips the
        packing calls) on the plan device's current stream (also what graph capture records).  The device guard makes the
        launches land on the GPU that owns the plan's buffers whatever the caller's current device is."""
        with torch.cuda.device(self.device):
            if self._ntags == 1 or os.environ.get("FT_BRANCHES", "1") == "0":
                s = stream_ptr()
                for what, fn, args, _keep in self.calls[start:]:
                    if fn is not None:
                        _lib.check(fn(*args, s), what)
                return
            streams = {0: torch.cuda.current_stream()}
            for i, (what, fn, args, _keep) in enumerate(self.calls):
                if i < start:
                    continue
                if fn is None:
                    child, parent = args
                    if child not in self._streams:
                        self._streams[child] = torch.cuda.Stream(device=self.device)
                    streams[child] = self._streams[child]
                    src, dst = (streams[parent], streams[child]) if what == "fork" else (streams[child], streams[parent])
                    ev = torch.cuda.Event()
                    ev.record(src)
                    dst.wait_event(ev)
                    continue
                _lib.check(fn(*args, ctypes.c_void_p(streams[self._tags.get(i, 0)].cuda_stream)), what)

    def time_calls(self, reps=3):
        """Per-call device time in ms (CUDA events on the launching stream, averaged over `reps` un-graphed passes).
        Event pairs around single launches include a few us of launch gap."""
        with torch.cuda.device(self.device):
            s = stream_ptr()
            n = len(self.calls)
            tot = [0.0] * n
            for _ in range(reps):
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
                ev[0].record()
                for i, (what, fn, args, _keep) in enumerate(self.calls):
                    if fn is not None:
                        _lib.check(fn(*args, s), what)
                    ev[i + 1].record()
                torch.cuda.synchronize()
                for i in range(n):
                    tot[i] += ev[i].elapsed_time(ev[i + 1])
        return [t / reps for t in tot]

    def autotune(self, launches=4, hint=None):
        """Pick the tile shape of every tensor-core convolution of the plan by measurement on this GPU and this batch shape:
        candidates = tile widths {64, 128, 256 (pairs only)} x {single CTA, CTA pair} x {1, 2 epilogue staging buffers}, and two
        CTAs per SM for 64-wide pair tiles; each candidate is captured as a small CUDA graph of
        `launches` back-to-back launches on the plan's own buffers and timed with CUDA events.  Which shape wins depends on
        how many tiles a layer has against the 148 SMs (wave quantisation) -- layer3/4 of a 64-crop batch, every layer of a
        2-pair strong-scaling shard -- which a static rule cannot see.  Layers with fewer tiles than SMs and a long K also try
        split-K (ft_conv_desc.splitk = 2..8).  Results are bit-identical across tile shapes (the K order of every accumulation
        chain is the same) and deterministic under split-K (parts are added in a fixed order; the chains are cut at other
        places, so those layers differ from the unsplit result by fp32 rounding, ~1e-7 relative); buffers hold whatever the last
        run left, timing does not depend on it.  FT_SPLITK=0 keeps split-K out of the candidates."""
        if self.impl != _lib.IMPL_TCGEN05 or torch.device(self.device).type != "cuda":
            return self
        fn = _lib.lib().ft_conv2d_fprop
        if hint is not None:
            # `hint` = the `tuned` record of a plan with the same layers and shapes (e.g. the other lane of a TrackingPipeline):
            # adopt its choices instead of measuring again
            return self._apply_tuned(hint)
        cache_path, sig = os.environ.get("FT_AUTOTUNE_CACHE"), None
        if cache_path:
            # choices measured by an earlier process on the same GPU model, kernel sources and plan (tools/dev/jobs/evidence.sh tunes
            # once outside the profiler and lets the runs under ncu load the result, so they profile the production tiling)
            sig = self.signature()
            try:
                with open(cache_path) as f:
                    rec = json.load(f).get(sig)
            except (OSError, ValueError):
                rec = None
            if rec is not None:
                return self._apply_tuned({int(i): tuple(v) for i, v in rec.items()})
        sms = torch.cuda.get_device_properties(self.device).multi_processor_count
        with torch.cuda.device(self.device):
            # one split-K workspace per branch of the plan (the layers of a branch run one after the other on one stream; branches
            # run concurrently and must not share it), sized for the branch's largest candidate
            need = {}
            plans = {}
            for i, (layer, x, out, out_f32, res) in sorted(self._conv_args.items()):
                plans[i] = []
                if layer.k_mode == _lib.K_PACKED or layer.block_n < 64 or os.environ.get("FT_SPLITK", "1") == "0":
                    continue
                for bn in (64, 128):
                    if layer.cout_pad % bn or bn > max(layer.block_n, 64):
                        continue
                    kb_total = layer.k_pad // BK + (-(-bn // BK) if layer.residual else 0)
                    for pair in (0, 1):
                        items, cta_tiles = layer.tile_counts(x.n, x.h, x.w, bn, pair)
                        units = sms // 2 if pair else sms
                        if kb_total < 8 or items >= 2.5 * units:
                            continue                         # enough tiles to fill the GPU, or too little K to cut
                        for sk in (2, 3, 4, 6, 8):
                            if sk * 4 > kb_total or items * sk > 6 * units:
                                continue
                            plans[i].append((bn, pair, 1, sk))
                            w, c = need.get(self._tags.get(i, 0), (0, 0))
                            need[self._tags.get(i, 0)] = (max(w, cta_tiles * sk * bn * 128 * 4), max(c, cta_tiles))
            self._sk_ws = {tag: (torch.empty(w, dtype=torch.uint8, device=self.device), torch.zeros(c, dtype=torch.int32, device=self.device))
                           for tag, (w, c) in need.items()}
            self.keep.append(self._sk_ws)
            for i, (layer, x, out, out_f32, res) in sorted(self._conv_args.items()):
                base = self.calls[i][3]
                cands = list(plans[i])
                for bn in (64, 128, 256):
                    if layer.cout_pad % bn or layer.block_n < 64 or (bn > max(layer.block_n, 64) and not (bn == 256 and layer.cout >= 256)):
                        continue
                    for pair in (0, 1):
                        if (pair and layer.k_mode == _lib.K_PACKED) or (bn == 256 and not pair):
                            continue
                        cands.append((bn, pair, 1))
                        if layer.k_mode != _lib.K_PACKED and out is not None and bn <= 128:
                            cands.append((bn, pair, 2))      # ping-pong epilogue staging (ft_conv_desc.epilogue_buffers)
                        if bn == 64 and pair and os.environ.get("FT_OCC2", "1") != "0":
                            cands.append((bn, pair, 1, 0, 2))    # two CTAs per SM (ft_conv_desc.ctas_per_sm)
                        if pair and bn <= 128 and layer.k_mode == _lib.K_TAP64 and (layer.cout_pad // bn) % 2 == 0 and os.environ.get("FT_QUAD", "0") != "0":
                            cands.append((bn, 2, 1))             # two CTA pairs per cluster sharing the A tile by TMA multicast (cta_pair = 2);
                                                                 # measured slower on every layer (profiles/r02m_autotune_log.txt), so opt-in: FT_QUAD=1
                if len(cands) < 2:
                    continue
                best = t_dflt = None
                dflt = tuple(layer.choose_tiling(x.n, x.h, x.w, self.precision)) + (1, 0, 1)
                sk = 0
                for cand in cands:
                    bn, pair, nbuf = cand[:3]
                    sk = cand[3] if len(cand) > 3 else 0
                    occ = cand[4] if len(cand) > 4 else 1
                    d = layer.desc(x, out, out_f32, res, self.precision, self.impl, tiling=(bn, pair, nbuf, sk, occ), splitk_ws=getattr(self, "_sk_ws", {}).get(self._tags.get(i, 0)))
                    d.tile_order = base.tile_order
                    try:
                        s = stream_ptr()
                        _lib.check(fn(ctypes.byref(d), s), "conv2d(autotune)")
                        torch.cuda.synchronize()
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g):
                            s = stream_ptr()
                            for _ in range(launches):
                                _lib.check(fn(ctypes.byref(d), s), "conv2d(autotune)")
                        g.replay()
                        torch.cuda.synchronize()
                        ts = []
                        for _ in range(3):
                            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                            e0.record()
                            g.replay()
                            e1.record()
                            torch.cuda.synchronize()
                            ts.append(e0.elapsed_time(e1) / launches)
                        del g
                        t = min(ts)
                    except RuntimeError:
                        continue
                    if os.environ.get("FT_AUTOTUNE_LOG"):
                        print("autotune call %3d %s k%d cin %4d cout %4d in %dx%dx%d : bn %3d pair %d nbuf %d splitk %d occ %d -> %7.1f us"
                              % (i, "deconv" if layer.kind else "conv", layer.kh, layer.cin, layer.cout, x.n, x.h, x.w, bn, pair, nbuf, sk, occ, t * 1e3), flush=True)
                    if best is None or t < best[0]:
                        best = (t, bn, pair, nbuf, sk, occ, d)
                    if (bn, pair, nbuf, sk, occ) == dflt:
                        t_dflt = (t, bn, pair, nbuf, sk, occ, d)
                if best is not None and t_dflt is not None and best[0] > t_dflt[0] * (1.0 - AUTOTUNE_MARGIN):
                    best = t_dflt                            # within the measurement noise of the static choice: keep that
                if best is not None:
                    t, bn, pair, nbuf, sk, occ, d = best
                    self.calls[i] = ("conv2d", fn, (ctypes.byref(d),), d)
                    self.tuned[i] = (bn, pair, t * 1e3, nbuf, sk, occ)
        self.graph = None
        if cache_path:
            try:
                with open(cache_path) as f:
                    cache = json.load(f)
            except (OSError, ValueError):
                cache = {}
            cache[sig] = {str(i): list(v) for i, v in self.tuned.items()}
            with open(cache_path, "w") as f:
                json.dump(cache, f, indent=0)
        return self

    def signature(self):
        """Identifies (GPU model, conv kernel sources, precision, every convolution's layer and input shape) for FT_AUTOTUNE_CACHE."""
        from . import build as ft_build
        h = hashlib.sha256((torch.cuda.get_device_name(self.device) + ft_build.kernel_hash(ft_build.CONV_SOURCES) + str(self.precision)).encode())
        for i, (layer, x, out, out_f32, res) in sorted(self._conv_args.items()):
            h.update(repr((i, layer.kind, layer.kh, layer.kw, layer.cin, layer.cout, layer.stride, layer.pad, layer.k_mode, layer.residual, x.n, x.h, x.w,
                           out is None, self._tags.get(i, 0))).encode())
        return h.hexdigest()

    def _apply_tuned(self, choices):
        """Rebuild the convolution descriptors from `choices` = {call index: (block_n, cta_pair, us, epilogue_buffers, splitk, ctas_per_sm)}
        (a `tuned` record); split-K choices get one workspace per branch of the plan, sized for the branch's largest user."""
        fn = _lib.lib().ft_conv2d_fprop
        with torch.cuda.device(self.device):
            need = {}
            for i, (bn, pair, us, nbuf, sk, occ) in choices.items():
                if sk > 1:
                    layer, x, out, out_f32, res = self._conv_args[i]
                    _items, cta_tiles = layer.tile_counts(x.n, x.h, x.w, bn, pair)
                    w, c = need.get(self._tags.get(i, 0), (0, 0))
                    need[self._tags.get(i, 0)] = (max(w, cta_tiles * sk * bn * 128 * 4), max(c, cta_tiles))
            self._sk_ws = {tag: (torch.empty(w, dtype=torch.uint8, device=self.device), torch.zeros(c, dtype=torch.int32, device=self.device))
                           for tag, (w, c) in need.items()}
            self.keep.append(self._sk_ws)
            for i, (bn, pair, us, nbuf, sk, occ) in choices.items():
                layer, x, out, out_f32, res = self._conv_args[i]
                d = layer.desc(x, out, out_f32, res, self.precision, self.impl, tiling=(bn, pair, nbuf, sk, occ), splitk_ws=self._sk_ws.get(self._tags.get(i, 0)))
                d.tile_order = self.calls[i][3].tile_order
                self.calls[i] = ("conv2d", fn, (ctypes.byref(d),), d)
                self.tuned[i] = (bn, pair, us, nbuf, sk, occ)
        self.graph = None
        return self

    def time_calls_graphed(self, reps=5):
        """Per-call device time in ms measured INSIDE a CUDA-graph replay: the plan is captured with an external
        (cudaEventRecordExternal) timing event between consecutive calls, replayed `reps` times, and the event pairs are read
        back -- kernels run back to back as in production, no host launch gaps and no per-launch tensor-map encoding in the
        intervals (time_calls() has both).  The event nodes sit between the kernels, so programmatic dependent launch cannot
        overlap across them: a call's figure here is its full, un-overlapped duration."""
        with torch.cuda.device(self.device):
            n = len(self.calls)
            self.enqueue()                                        # warm-up outside the capture (timed below in linear order, branches serialised)
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True, external=True) for _ in range(n + 1)]
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                s = stream_ptr()
                ev[0].record()
                for i, (what, fn, args, _keep) in enumerate(self.calls):
                    if fn is not None:
                        _lib.check(fn(*args, s), what)
                    ev[i + 1].record()
            tot = [0.0] * n
            g.replay()
            torch.cuda.synchronize()
            for _ in range(reps):
                g.replay()
                torch.cuda.synchronize()
                for i in range(n):
                    tot[i] += ev[i].elapsed_time(ev[i + 1])
            del g
        return [t / reps for t in tot]

    def run(self):
        if self.graph is not None:
            with torch.cuda.device(self.device):
                self.graph.replay()
            return
        self.enqueue()

    def capture(self):
        """Capture the call list into a CUDA graph (one launch per forward instead of one per layer)."""
        with torch.cuda.device(self.device):
            self.run()                      # warm-up: sets function attributes, faults in code
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.enqueue()
            self.graph = g
        return self
