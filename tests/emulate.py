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


def space_to_depth16(x):
    """[B,3,H,W] -> [B,H/2,W/2,16] NHWC: channel (py*2+px)*4 + c = x[:, c, 2Y+py, 2X+px] (ft_pack_image_s2d)."""
    B, C, H, W = x.shape
    out = torch.zeros(B, H // 2, W // 2, 16, dtype=x.dtype)
    for py in range(2):
        for px in range(2):
            out[..., (py * 2 + px) * 4:(py * 2 + px) * 4 + C] = x[:, :, py::2, px::2].permute(0, 2, 3, 1)
    return out


def emulate_pose(model, x):
    """Walk DeconvResnet exactly like DeconvResnet._build_plan does, on the CPU."""
    L = model._pack(None, keep_host=True)
    if L["stem"].cin_p == 16:          # space-to-depth stem: mirror of ft_pack_image_s2d
        t = emulate_layer(L["stem"], space_to_depth16(x))
    else:
        t = emulate_layer(L["stem"], nhwc(x, 8))
    t = torch.nn.functional.max_pool2d(t.permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
    for blk in L["blocks"]:
        a = emulate_layer(blk["c2"], emulate_layer(blk["c1"], t))
        r = t if blk["down"] is None else emulate_layer(blk["down"], t)
        t = emulate_layer(blk["c3"], a, res=r)
    for dl in L["deconv"]:
        t = emulate_layer(dl, t)
    return emulate_layer(L["head"], t).permute(0, 3, 1, 2).contiguous()


def emulate_corr_banded(in1, in2):
    """Mirror of corr_tc_kernel's decomposition (csrc/ft_corr.cu): four parity classes, tiles of 4 rows x 32 columns of
    in1', chunks of 4 rows x 32 columns of in2', all-pairs GEMM per (tile, chunk), 21-wide band extraction, explicit
    zeros for displacement rows outside the image.  in1, in2: [B,C,H,W] -> [B,441,H,W] (float64 arithmetic)."""
    B, C, H, W = in1.shape
    assert (W + 1) // 2 <= 32
    out = torch.full((B, 441, H, W), float("nan"), dtype=torch.float64)
    D, ROWS, COLS = 10, 4, 32
    H2 = (H + 1) // 2
    for n in range(B):
        for cls in range(4):
            py, px = cls >> 1, cls & 1
            a_img = in1[n, :, py::2, px::2].double()        # [C, H2c, W2c]
            b_img = in2[n, :, py::2, px::2].double()
            H2c, W2c = (H - py + 1) // 2, (W - px + 1) // 2

            def box(img, r0):                               # TMA box {C, 32 cols from 0, 4 rows from r0}: zero fill outside
                t = torch.zeros(C, ROWS, COLS, dtype=torch.float64)
                for r in range(ROWS):
                    if 0 <= r0 + r < H2c and W2c > 0:
                        t[:, r, :W2c] = img[:, r0 + r]
                return t.reshape(C, ROWS * COLS)            # column index = r*32 + b

            for at in range((H2 + ROWS - 1) // ROWS):
                a0 = at * ROWS
                r_lo, r_end = max(0, a0 - D), min(H2c, a0 + ROWS + D)
                chunks = (r_end - r_lo + ROWS - 1) // ROWS if r_end > r_lo else 0
                A = box(a_img, a0)
                written = set()
                for q in range(ROWS):                       # the zero loop of the epilogue warps
                    a = a0 + q
                    for tj in range(-D, D + 1):
                        row = a + tj
                        if 0 <= row < H2c:
                            continue
                        written.add((q, tj))
                        y = 2 * a + py
                        for lane in range(COLS):
                            x = 2 * lane + px
                            if y < H and x < W:
                                out[n, (tj + D) * 21:(tj + D) * 21 + 21, y, x] = 0.0
                for j in range(chunks):
                    r0 = r_lo + j * ROWS
                    Dm = A.t() @ box(b_img, r0)             # [128 in1' pixels, 128 in2' pixels]
                    for q in range(ROWS):
                        a = a0 + q
                        y = 2 * a + py
                        for rr in range(ROWS):
                            row = r0 + rr
                            tj = row - a
                            if not (-D <= tj <= D and row < H2c):
                                continue
                            assert (q, tj) not in written
                            written.add((q, tj))
                            for lane in range(COLS):
                                x = 2 * lane + px
                                if not (y < H and x < W):
                                    continue
                                for ti in range(-D, D + 1):
                                    b2 = lane + ti
                                    v = Dm[q * COLS + lane, rr * COLS + b2] if 0 <= b2 < COLS else 0.0
                                    out[n, (tj + D) * 21 + ti + D, y, x] = v / C
                assert len(written) == ROWS * 21
    return out


def corr_pair_unit_coverage(H, parts=3):
    """Mirror of corr_pair_unit (csrc/ft_corr.cu): for a parity class with H2c rows, the (in1' row, displacement row) pairs
    every CTA of every pair unit handles.  Returns {(a, tj): count} over all units -- each in-image pair must appear once."""
    D, ROWS = 10, 4
    out = {}
    for py in (0, 1):
        H2c = (H - py + 1) // 2
        a_tiles = ((H + 1) // 2 + ROWS - 1) // ROWS
        ap_tiles = (a_tiles + 1) // 2
        cover = {}
        for ap in range(ap_tiles):
            A0 = ap * 2 * ROWS
            r_lo, r_end = max(0, A0 - D), min(H2c, A0 + 2 * ROWS + D)
            chunks = (r_end - r_lo + ROWS - 1) // ROWS if r_end > r_lo else 0
            per = (chunks + parts - 1) // parts
            for part in range(parts):
                c_lo = min(chunks, part * per)
                nch = min(chunks, c_lo + per) - c_lo
                for rank in (0, 1):
                    for q in range(ROWS):
                        a = A0 + rank * ROWS + q
                        if a >= H2c:
                            continue                                   # rows beyond the image: nothing is stored
                        if part == 0:                                  # the zero loop
                            for tj in range(-D, D + 1):
                                if not (0 <= a + tj < H2c):
                                    cover[(a, tj)] = cover.get((a, tj), 0) + 1
                        for j in range(nch):
                            r0 = r_lo + (c_lo + j) * ROWS
                            for rr in range(ROWS):
                                row = r0 + rr
                                tj = row - a
                                if -D <= tj <= D and row < H2c:
                                    cover[(a, tj)] = cover.get((a, tj), 0) + 1
        out[py] = (H2c, cover)
    return out


# ---- mirrors of the conv kernel's tile bookkeeping (csrc/ft_conv.cu: make_fastdiv / fdiv / tile_coord / choose_box) ----
def make_fastdiv(d):
    l = 0
    while (1 << l) < d:
        l += 1
    m = (((1 << 32) * ((1 << l) - d)) // d + 1) & 0xffffffff
    return m, l, d


def fdiv(n, f):
    m, l, _ = f
    return ((((n * m) >> 32) + n) & 0xffffffffffffffff) >> l


def choose_box(qw, qh, n):
    best, box = -1, None
    for lw in range(7, -1, -1):
        for lh in range(7 - lw, -1, -1):
            w, h = 1 << lw, 1 << lh
            nn = 128 // (w * h)
            tiles = -(-qw // w) * -(-qh // h) * -(-n // nn)
            if best < 0 or tiles < best:
                best, box = tiles, (w, h, nn)
    return box


def conv_tile_coverage(qw, qh, n, n_tiles, n_phase, pair, rev=False):
    """Every (phase, m-tile, n-tile) the persistent CTAs of conv_tc_kernel visit, decoded like tile_coord does (multiply-high
    division, CTA pairs taking two consecutive m-tiles).  Returns (visited list, m_tiles, pixel coverage count array)."""
    bw, bh, bn = choose_box(qw, qh, n)
    tiles_w, tiles_h = -(-qw // bw), -(-qh // bh)
    m_tiles = tiles_w * tiles_h * -(-n // bn)
    m_units = (m_tiles + 1) // 2 if pair else m_tiles
    num_tiles = n_phase * m_units * n_tiles
    f_n, f_m, f_w, f_h = make_fastdiv(n_tiles), make_fastdiv(m_units), make_fastdiv(tiles_w), make_fastdiv(tiles_h)
    visited = []
    cover = np.zeros((n_phase, n_tiles, n, qh, qw), np.int32)
    for tile0 in range(num_tiles):
        for rank in ((0, 1) if pair else (0,)):
            tile = num_tiles - 1 - tile0 if rev else tile0
            r = fdiv(tile, f_n)
            nt = tile - r * n_tiles
            phase = fdiv(r, f_m)
            mu = r - phase * m_units
            mt = mu * 2 + rank if pair else mu
            assert (r, nt, phase, mu) == (tile // n_tiles, tile % n_tiles, (tile // n_tiles) // m_units, (tile // n_tiles) % m_units)
            r2 = fdiv(mt, f_w)
            tw = mt - r2 * tiles_w
            r3 = fdiv(r2, f_h)
            qx0, qy0, n0 = tw * bw, (r2 - r3 * tiles_h) * bh, r3 * bn
            visited.append((phase, mt, nt))
            if mt >= m_tiles:
                assert n0 >= n                     # the odd tail of a pair: an all-padding box
                continue
            cover[phase, nt, n0:n0 + bn, qy0:qy0 + bh, qx0:qx0 + bw] += 1
    return visited, m_tiles, cover
