// ft_conv.cu -- fused convolution / k4s2p1 transposed convolution as an implicit GEMM on the
// Blackwell tensor cores (tcgen05.mma, accumulators in TMEM, operands fed by TMA), sm_100a only.
//
//   out[m, co] = act( (sum_k A[m,k] * Wt[co,k]) * scale[co] + shift[co] )        (residual: see below)
//     m  = output pixel.  A tile is 128 pixels = a box of bw x bh x bn (columns x rows x images) of the
//          output grid, bw*bh*bn = 128, chosen per layer so the 8x6 maps of layer4 fill tiles as well as
//          the 64x48 ones.
//     k  = (tap, 64-channel block): implicit im2col -- for every tap the A tile is ONE TMA box load of
//          the NHWC activation tensor at the tap's offset; image borders (padding) and channel tails are
//          zero-filled by the TMA unit, so no thread ever computes an address.
//          Stride-2 convolutions read through four "parity views" of the input (tensor maps whose pixel
//          strides are doubled), transposed convolutions write through four parity views of the output.
//   residual (bottleneck conv3): folded into the GEMM as extra k-blocks  acc += R[m, 64-block] * I*f,
//          R read through the same TMA pipeline as A, so the epilogue never waits on a global load
//          (host side folds BatchNorm's scale into the weights for these layers; engine.py).
//
// CTA = 10 warps (14 for the gather variant), persistent over output tiles (static round-robin):
//   warps 0-7  epilogue : two groups of four warps (one warp per TMEM lane quarter); group g takes the accumulator
//                         slabs g, g+2, ... of 16/32/64 columns: tcgen05.ld (lane = pixel row), scale/shift
//                         (folded BN or bias), ReLU/LeakyReLU, re-split to fp16 hi/lo, st.shared into the group's
//                         swizzled staging slab, then ONE TMA store per plane writes the slab to the NHWC output
//                         (coalesced, clipped at tensor edges).  The TMEM accumulator is double buffered, so the
//                         epilogue of tile i overlaps the main loop of tile i+1.  (fp32 NCHW network outputs are
//                         written straight from registers.)
//   warp 8     MMA      : one thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16) four
//                         times per 64-wide k-block -- three times that in split mode:
//                         acc += A_hi*W_hi + A_lo*W_hi + A_hi*W_lo  (fp16 hi/lo pairs carry 22 significant
//                         bits, which is what the 1e-3 heat-map parity gate needs).
//   warp 9     TMA      : one thread: per k-block one cp.async.bulk.tensor.4d per activation plane plus one
//                         cp.async.bulk of the pre-swizzled weight tile, all completing on the stage's mbarrier.
//   warps 10-13 gather  : only in the GATHER variant (Cin < 64: the 7x7 / 3x3 stems, where a 64-channel box per
//                         tap would be mostly padding): im2col by 16-byte cp.async with K packed across taps.
//   mbarriers: full[s] (TMA bytes -> MMA), empty[s] (tcgen05.commit -> producer),
//              acc_full[a] (tcgen05.commit -> epilogue), acc_empty[a] (epilogue -> MMA).
//
// The SIMT kernel at the bottom computes the same thing with fp32 FMAs on the same data formats; it is
// the on-GPU cross-check used by the tests (FT_IMPL_SIMT), not a fallback: ft_conv2d_fprop never
// switches implementation on its own.
#include <cuda.h>
#include <stdlib.h>

#include "ft_tc.cuh"

namespace ft {

// Division by a runtime constant as multiply-high + shift (Granlund-Montgomery round-up magic; exact for n < 2^31):
// tile_coord runs in every role once per tile, and four hardware-less integer divisions were ~15 % of the epilogue's
// instructions on the small-K layers.
struct FastDiv {
    uint32_t m, l;
    int d;
};
static inline FastDiv make_fastdiv(int d) {
    FastDiv f;
    f.d = d;
    f.l = 0;
    while ((1u << f.l) < (uint32_t)d) ++f.l;                       // ceil(log2 d)
    f.m = (uint32_t)((((uint64_t)1 << 32) * (((uint64_t)1 << f.l) - (uint64_t)d)) / (uint64_t)d + 1);
    return f;
}
__device__ __forceinline__ int fdiv(int n, const FastDiv& f) { return (int)((__umulhi((uint32_t)n, f.m) + (uint32_t)n) >> f.l); }

struct alignas(64) ConvK {
    CUtensorMap in_map[4][2];    // [input parity view][hi, lo]   (stride 1: only view 0)
    CUtensorMap out_map[4][2];   // [output phase][hi, lo]        (plain conv: only phase 0)
    CUtensorMap res_map[2];      // residual as an A-operand source
    CUtensorMap in_map_h[4][2];  // QUAD: the same views with HALF pixel boxes (64 pixels): each CTA loads one half and multicasts it
    const __half* in_hi;
    const __half* in_lo;
    __half* out_hi;
    __half* out_lo;
    float* out_f32;
    const __half* res_hi;
    const __half* res_lo;
    const uint8_t* w_tile;
    const float* w_gemm;
    const float* scale;
    const float* shift;
    int N, H, W, Wp, xoff, in_ld, in_coff, cin_p, cin_k, cpt;   // Wp/xoff: physical row pitch (pixels) and left padding of the input
    int kw, pad, row_pad, row_mode;                                   // FT_K_ROW: one k-block per kernel row
    int Qh, Qw, Ho, Wo, out_ld, out_coff, cout, cout_p8, cout_pad;
    int res_ld, res_coff;
    int in_stride, out_mul;
    int n_phase, taps, cb, kb_main, kb_total, k_pad, chunk;
    FastDiv fd_n_tiles, fd_m_units, fd_tiles_w, fd_tiles_h;
    int rev;        // tile_order: walk the tile list from the end (see ft_conv_desc.tile_order)
    int dbg_skip;   // bring-up/profiling knob (FT_DEV_SKIP): 1 = no A loads, 2 = no weight loads, 4 = no MMAs, 8 = no TMA stores; split-K: 16 = no partial-sum
                    // stores, 32 = every thread fences (the old protocol), 128 = the last arriver reads part 0 only (results are garbage)
    int M, m_tiles, n_tiles, num_tiles;
    int n_tiles_w;               // n-tiles of the weight image (= n_tiles, or 2 * n_tiles for QUAD, whose work items span two n-tiles)
    int half_dx, half_dy, half_dn;   // QUAD: offset of the second half box inside a tile's pixel box
    // split-K (ft_conv_desc.splitk): a work item is (tile, k-split); items = num_tiles * splits, item -> split = item / num_tiles
    int splits, kb_split, num_items;
    FastDiv fd_num_tiles;
    float* sk_ws;            // partial accumulators [cta tile][split][BN / 4 column quads][128 rows][4] fp32
    unsigned* sk_cnt;        // arrivals per cta tile (zero between launches: the last arriver resets its slot)
    int bw, bh, bn, lg_bw, lg_bh, tiles_w, tiles_h;
    int act;
    float slope;
    float act_mul;   // tensor-core epilogue: out = max(v, min(act_mul * v, act_cap)); act_cap = 0 for ReLU (so that -inf -> 0, not NaN), +inf otherwise
    float act_cap;
    int8_t tap_dy[4][64];
    int8_t tap_dx[4][64];
    int8_t ph_oy[4];
    int8_t ph_ox[4];
};

// ------------------------------------------------------------------------------------------------
// Tile geometry shared by host and device
// ------------------------------------------------------------------------------------------------
constexpr int BM = 128;              // pixels per tile (UMMA M)
constexpr int BK = 64;               // fp16 per k-block = one 128-byte swizzle row
constexpr int A_TILE = BM * BK * 2;  // 16 KiB
constexpr int EPI_WARPS = 8;            // two groups of 4 warps (one per TMEM lane quarter); group g takes slabs g, g+2, ...
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int MMA_WARP = EPI_WARPS, TMA_WARP = EPI_WARPS + 1, GATHER_WARP0 = EPI_WARPS + 2;
constexpr int THREADS_TMA = (EPI_WARPS + 2) * 32, THREADS_GATHER = (EPI_WARPS + 6) * 32;
constexpr int ACC_CHUNK = 4;         // split mode: k-blocks (of 64) per tensor-core accumulation chain
constexpr int SMEM_LIMIT = 232448;   // 227 KiB opt-in maximum per CTA

template <int BN, bool SPLIT, bool PAIR = false, int NBUF = 1, int OCC = 1>
struct Geo {
    static constexpr int PLANES = SPLIT ? 2 : 1;
    static constexpr int B_TILE = BN * BK * 2;                     // whole weight k-block of the n-tile (one plane)
    static constexpr int B_LOCAL = PAIR ? B_TILE / 2 : B_TILE;     // pair: each CTA stages its half of the BN rows
    static constexpr int STAGE = PLANES * (A_TILE + B_LOCAL);
    // epilogue staging: per warp group NBUF slabs of (SLAB output channels x 128 pixels x 2 planes).  NBUF = 2 lets a group fill
    // one slab while the TMA store of its previous slab is still reading the other (the small-K 1x1 layers are bound by the
    // epilogue's store hand-off chain, not by the tensor pipe); it costs 32 KiB of shared memory, i.e. a pipeline stage.
    // OCC = 2: two CTAs (two CTA pairs per TPC) share an SM -- half the shared memory each (narrow 16-channel slabs, 2 stages), so one
    // CTA's epilogue / pipeline-fill latency runs under the other's main loop.  For the layers whose time is the per-tile epilogue
    // latency chain, not the tensor pipe or HBM: the stems, 64-channel 3x3s and deconvs (epilogue throughput ~2.2 outputs/clk/SM).
    static constexpr int SLAB = (BN >= 64 && OCC == 1) ? 32 : 16;
    static constexpr int SLAB_PLANE = BM * SLAB * 2;
    static constexpr int GROUP_STAGING = NBUF * 2 * SLAB_PLANE;
    static constexpr int STAGING = 2 * GROUP_STAGING;
    static constexpr int SS_BYTES = 2 * 2 * BN * 4;        // the n-tile's scale | shift, double buffered across tiles (see the epilogue)
    static constexpr int FIXED = 1024 /*align slack*/ + 512 /*barriers*/ + SS_BYTES + 16 /*split-K flag*/ + STAGING;
    static_assert(STAGE % 1024 == 0, "stages must keep the 1024-byte swizzle alignment");
    static constexpr int STAGES_FIT = ((OCC == 2 ? SMEM_LIMIT / 2 - 1024 : SMEM_LIMIT) - FIXED) / STAGE;   // OCC 2: 1 KiB per CTA is reserved by the system
    static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;
    static constexpr int LAG = STAGES >= 3 ? 2 : 1;       // gather variant: cp.async groups a producer keeps in flight
    static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
    static constexpr int SMEM = STAGES * STAGE + FIXED;
    static_assert(STAGES >= 2, "tile too large for shared memory");
};

struct TileCoord {
    int phase, mt, nt;
    int qx0, qy0, n0;    // TMA variant: origin of the pixel box
};
template <bool GATHER, bool PAIR = false, bool QUAD = false>
__device__ __forceinline__ TileCoord tile_coord(const ConvK& p, int tile, int cta_rank = 0, int pair_id = 0) {
    TileCoord t;
    if (p.rev) tile = p.num_tiles - 1 - tile;
    const int r = fdiv(tile, p.fd_n_tiles);
    t.nt = tile - r * p.n_tiles;           // consecutive CTAs share the A tile (L2 reuse), differ in weights
    if (QUAD) t.nt = 2 * t.nt + pair_id;   // a quad's two CTA pairs take two neighbouring n-tiles of the same pixels
    // fd_m_units divides by the m-tiles (single CTA) or the m-tile pairs (PAIR: two consecutive m-tiles, one per CTA, x one n-tile)
    t.phase = fdiv(r, p.fd_m_units);
    const int mu = r - t.phase * p.fd_m_units.d;
    t.mt = PAIR ? mu * 2 + cta_rank : mu;  // pair: may be == m_tiles for the odd tail: an all-padding box (zero loads, clipped stores)
    if (GATHER) {
        t.qx0 = t.mt * BM, t.qy0 = 0, t.n0 = 0;   // linear pixel index; the output map is [C, M]
    } else {
        const int r2 = fdiv(t.mt, p.fd_tiles_w), tw = t.mt - r2 * p.tiles_w;
        const int r3 = fdiv(r2, p.fd_tiles_h);
        t.qx0 = tw * p.bw;
        t.qy0 = (r2 - r3 * p.tiles_h) * p.bh;
        t.n0 = r3 * p.bn;
    }
    return t;
}

// Work item -> (tile, k-block range).  Without split-K an item is a whole tile.  With it, the `splits` parts of a tile are
// `num_tiles` items apart, so they run on different CTAs at the same time and a layer with fewer tiles than SMs still fills the GPU.
struct Item {
    int tile, sp, kb_lo, kb_hi;
};
__device__ __forceinline__ Item work_item(const ConvK& p, int item) {
    Item w;
    w.sp = p.splits > 1 ? fdiv(item, p.fd_num_tiles) : 0;
    w.tile = item - w.sp * p.num_tiles;
    w.kb_lo = w.sp * p.kb_split;
    w.kb_hi = min(w.kb_lo + p.kb_split, p.kb_total);
    return w;
}

// PAIR: two CTAs of a cluster (the two SMs of a TPC) work on one 256-pixel x BN tile with tcgen05.mma.cta_group::2:
// each CTA stages its own 128 pixel rows of A and HALF of the BN weight rows, the tensor cores of both SMs read both
// halves, and each CTA's TMEM receives its 128 rows of the result.  Per MMA cycle a CTA then needs half the weight
// bytes -- the main loop of the single-CTA variant is bound by shared-memory fill bandwidth (~48 B/clk/SM measured),
// not by the tensor pipe.  CTA rank 0 (the leader) issues every MMA; the peer's MMA warp only relays "my stage is
// full" to the leader.
// QUAD (PAIR only): clusters of FOUR CTAs = two CTA pairs that work on the SAME 256 pixels and two neighbouring n-tiles.  The A
// operand (the activations: two thirds of a k-block's bytes) is identical for CTA r of both pairs, so each of the two loads one HALF
// of it (64 pixels) and TMA-multicasts it into both: a k-block costs 32 KiB of L2 -> SM traffic per SM instead of 48 KiB.  ncu
// shows the conv stack at 10-11 TB/s of L2 -> L1 read traffic on nearly every large layer: that, not the tensor pipe,
// is the bound this attacks.  A stage may be refilled only when BOTH pairs have consumed it (the refill lands in both CTAs), so the
// empty barriers count two commits, each multicast to all four CTAs.
template <int BN, bool SPLIT, bool GATHER, bool PAIR, int NBUF = 1, int OCC = 1, bool QUAD = false>
__global__ void __launch_bounds__(GATHER ? THREADS_GATHER : THREADS_TMA, OCC) conv_tc_kernel(const __grid_constant__ ConvK p) {
    static_assert(!(GATHER && PAIR), "the gather variant is single-CTA");
    static_assert(!QUAD || (PAIR && OCC == 1), "QUAD = two CTA pairs per cluster");
    static_assert(OCC == 1 || (OCC == 2 && !GATHER && NBUF == 1 && 2 * BN <= 256), "two CTAs per SM: TMA variant, half of TMEM each");
    using G = Geo<BN, SPLIT, PAIR, NBUF, OCC>;
    constexpr int CL = QUAD ? 4 : (PAIR ? 2 : 1);                                 // CTAs per cluster
    const int crank = PAIR ? (int)cluster_ctarank() : 0;
    const int cta_rank = crank & 1;                                               // rank inside the CTA pair
    const int pair_id = QUAD ? crank >> 1 : 0;
    const uint32_t leader_rank = (uint32_t)(pair_id * 2);                         // cluster rank of this pair's MMA-issuing CTA
    const int first_tile = (int)(blockIdx.x / CL);                                // a cluster owns a work item
    const int tile_step = (int)(gridDim.x / CL);
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t staging = smem_base + G::STAGES * G::STAGE;         // 1024-aligned (STAGE is a multiple of 1024)
    const uint32_t bar_base = staging + G::STAGING;
    // barrier slots (8 bytes each): full[STAGES] | empty[STAGES] | acc_full[2] | acc_empty[2] | tmem ptr | peer_full[STAGES]
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (G::STAGES + s); };
    auto accf_bar = [&](int a) { return bar_base + 8u * (2 * G::STAGES + a); };
    auto acce_bar = [&](int a) { return bar_base + 8u * (2 * G::STAGES + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * G::STAGES + 4);
    auto peer_full_bar = [&](int s) { return bar_base + 8u * (2 * G::STAGES + 5 + s); };   // leader only: the peer's stage s is full
    const uint32_t ss_base = bar_base + 512u;            // [2 tile parities][scale BN | shift BN] floats

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < G::STAGES; ++s) {
            mbar_init(full_bar(s), GATHER ? 128 + 1 : 1);   // (128 gather arrivals +) 1 arrive.expect_tx for the TMA bytes
            mbar_init(empty_bar(s), QUAD ? 2 : 1);          // one tcgen05.commit (QUAD: one per CTA pair, see above)
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(accf_bar(a), 1);
            mbar_init(acce_bar(a), PAIR ? 2 * EPI_THREADS : EPI_THREADS);   // pair: the epilogue threads of both CTAs release the stage
        }
        if (PAIR)
            for (int s = 0; s < G::STAGES; ++s) mbar_init(peer_full_bar(s), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)G::TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)G::TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();          // both CTAs' barriers exist before any cross-CTA arrive / multicast commit
    tc_fence_after();
    // PDL: this CTA is set up -- let the next kernel of the stream take SMs as they free up, then wait until the previous kernel's
    // results are complete and visible before the first global read / write of this one (all roles, see ft_tc.cuh)
    pdl_launch_dependents();
    pdl_wait();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == TMA_WARP) {
        // =========================== TMA producer ===========================
        // the whole warp runs the loop with warp-uniform values and one elected lane issues the copies: in single-lane code the
        // compiler wraps every UTMALDG / UTCHMMA in an ELECT + R2UR.BROADCAST sequence (~100 cycles each)
        if (!GATHER) {
            int it = 0;
            for (int item = first_tile; item < p.num_items; item += tile_step) {
                const Item wi = work_item(p, item);
                const TileCoord tc = tile_coord<false, PAIR, QUAD>(p, wi.tile, cta_rank, pair_id);
                // weight image: per (phase, n_tile, k_block) [hi | lo][BN rows][64]; a pair CTA copies rows [rank*BN/2, +BN/2) of each plane
                const uint8_t* wsrc = p.w_tile + ((size_t)(tc.phase * p.n_tiles_w + tc.nt) * p.kb_total) * (size_t)(2 * G::B_TILE) + (size_t)cta_rank * G::B_LOCAL;
                int tap = wi.kb_lo / p.cb, cblk = wi.kb_lo - tap * p.cb;          // (tap, channel block) of the first k-block of this item
                for (int kb = wi.kb_lo; kb < wi.kb_hi; ++kb, ++it) {
                    const int s = it % G::STAGES;
                    mbar_wait(empty_bar(s), ((it / G::STAGES) & 1) ^ 1);
                    const uint32_t stage = smem_base + s * G::STAGE;
                    const bool ldA = !(p.dbg_skip & 1), ldB = !(p.dbg_skip & 2);
                    const bool leader = elect_one();
                    if (leader) mbar_arrive_expect_tx(full_bar(s), (uint32_t)((ldA ? G::PLANES * A_TILE : 0) + (ldB ? G::PLANES * G::B_LOCAL : 0)));
                    if (!ldA || !leader) {
                    } else if (kb < p.kb_main) {
                        // FT_K_ROW: k-block `tap` is kernel row tap (all kw taps of it, contiguous in the x-padded input)
                        const int dy = p.row_mode ? tap - p.pad : p.tap_dy[tc.phase][tap], dx = p.row_mode ? 0 : p.tap_dx[tc.phase][tap];
                        int view = 0, cx = tc.qx0 + dx, cy = tc.qy0 + dy;
                        if (p.in_stride == 2) {      // input pixel 2q+d lives in parity view (d&1) at index q + (d - (d&1))/2
                            const int py = dy & 1, px = dx & 1;
                            view = py * 2 + px;
                            cx = tc.qx0 + ((dx - px) >> 1);
                            cy = tc.qy0 + ((dy - py) >> 1);
                        }
                        if (QUAD) {
                            // this CTA fetches half `pair_id` (64 pixels) of the A tile and multicasts it to itself and to its twin in the
                            // other pair (same rank inside the pair): both full barriers receive both halves' bytes
                            const uint16_t mask = (uint16_t)((1u << cta_rank) | (1u << (2 + cta_rank)));
                            const uint32_t dst = stage + (uint32_t)pair_id * (A_TILE / 2);
                            const int hx = cx + pair_id * p.half_dx, hy = cy + pair_id * p.half_dy, hn = tc.n0 + pair_id * p.half_dn;
                            tma_load_4d_mc(dst, &p.in_map_h[view][0], full_bar(s), cblk * BK, hx, hy, hn, mask);
                            if (SPLIT) tma_load_4d_mc(dst + A_TILE, &p.in_map_h[view][1], full_bar(s), cblk * BK, hx, hy, hn, mask);
                        } else {
                            tma_load_4d(stage, &p.in_map[view][0], full_bar(s), cblk * BK, cx, cy, tc.n0);
                            if (SPLIT) tma_load_4d(stage + A_TILE, &p.in_map[view][1], full_bar(s), cblk * BK, cx, cy, tc.n0);
                        }
                    } else {                          // residual k-block: 64 channels of the residual tensor at the output pixels
                        const int c0 = tc.nt * BN + (kb - p.kb_main) * BK;
                        tma_load_4d(stage, &p.res_map[0], full_bar(s), c0, tc.qx0, tc.qy0, tc.n0);
                        if (SPLIT) tma_load_4d(stage + A_TILE, &p.res_map[1], full_bar(s), c0, tc.qx0, tc.qy0, tc.n0);
                    }
                    if (kb < p.kb_main && ++cblk == p.cb) cblk = 0, ++tap;      // (tap, channel block) of the next k-block: every lane tracks it
                    if (!ldB || !leader) {
                    } else if (PAIR) {
                        bulk_g2s(stage + G::PLANES * A_TILE, wsrc + (size_t)kb * (2 * G::B_TILE), G::B_LOCAL, full_bar(s));
                        if (SPLIT) bulk_g2s(stage + G::PLANES * A_TILE + G::B_LOCAL, wsrc + (size_t)kb * (2 * G::B_TILE) + G::B_TILE, G::B_LOCAL, full_bar(s));
                    } else {
                        bulk_g2s(stage + G::PLANES * A_TILE, wsrc + (size_t)kb * (2 * G::B_TILE), G::PLANES * G::B_TILE, full_bar(s));
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= GATHER_WARP0) {
        // =========================== gather producers (GATHER variant only) ===========================
        if (GATHER) {
            const int ptid = threadIdx.x - GATHER_WARP0 * 32;
            const int j = ptid & 7;             // 16-byte chunk (8 channels) inside the 64-wide k-block
            const int r0 = ptid >> 3;           // rows r0 + 16*i
            const uint32_t row_off = (uint32_t)r0 * 128u + (uint32_t)((j ^ (r0 & 7)) << 4);   // 128B swizzle: chunk ^= row%8
            int it = 0;                         // k-block iterations issued (across tiles)
            int arrived = 0;                    // iterations whose full-barrier arrival has been made
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const TileCoord tc = tile_coord<true>(p, tile);
                int pix[8], iy0[8], ix0[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int m = tc.mt * BM + r0 + 16 * i;
                    if (m < p.M) {
                        const int qx = m % p.Qw, t2 = m / p.Qw;
                        const int qy = t2 % p.Qh, n = t2 / p.Qh;
                        iy0[i] = qy * p.in_stride;
                        ix0[i] = qx * p.in_stride;
                        pix[i] = (n * p.H + iy0[i]) * p.Wp + ix0[i] + p.xoff;
                    } else {
                        iy0[i] = -(1 << 20);    // fails every bounds test -> zero fill
                        ix0[i] = 0;
                        pix[i] = 0;
                    }
                }
                const uint8_t* wsrc = p.w_tile + ((size_t)(tc.phase * p.n_tiles + tc.nt) * p.kb_total) * (size_t)(2 * G::B_TILE);
                for (int kb = 0; kb < p.kb_total; ++kb, ++it) {
                    const int s = it % G::STAGES;
                    mbar_wait(empty_bar(s), ((it / G::STAGES) & 1) ^ 1);
                    const uint32_t stage = smem_base + s * G::STAGE;
                    if (ptid == 0) {
                        constexpr uint32_t wbytes = G::PLANES * G::B_TILE;
                        mbar_arrive_expect_tx(full_bar(s), wbytes);
                        bulk_g2s(stage + G::PLANES * A_TILE, wsrc + (size_t)kb * (2 * G::B_TILE), wbytes, full_bar(s));
                    }
                    const int g = kb * 8 + j;            // global 8-channel chunk index along K (K packed across taps)
                    const int tap = g / p.cpt;
                    const int c8 = g - tap * p.cpt;
                    const bool tap_ok = tap < p.taps;
                    const int dy = tap_ok ? p.tap_dy[tc.phase][tap] : 0;
                    const int dx = tap_ok ? p.tap_dx[tc.phase][tap] : 0;
                    const int doff = dy * p.Wp + dx;
                    const int coff = p.in_coff + c8 * 8;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int iy = iy0[i] + dy, ix = ix0[i] + dx;
                        const bool ok = tap_ok && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
                        const size_t off = ok ? ((size_t)(pix[i] + doff) * p.in_ld + coff) : 0;
                        const uint32_t dst = stage + row_off + (uint32_t)i * (16u * 128u);
                        cp_async16(dst, p.in_hi + off, ok ? 16u : 0u);
                        if (SPLIT) cp_async16(dst + A_TILE, p.in_lo + off, ok ? 16u : 0u);
                    }
                    cp_async_commit();
                    if (it - arrived >= G::LAG) {
                        cp_async_wait<G::LAG>();
                        fence_proxy_async();
                        mbar_arrive(full_bar(arrived % G::STAGES));
                        ++arrived;
                    }
                }
            }
            cp_async_wait<0>();
            fence_proxy_async();
            for (; arrived < it; ++arrived) mbar_arrive(full_bar(arrived % G::STAGES));
        }
    } else if (warp == MMA_WARP) {
        // =========================== MMA issuer ===========================
        if (cta_rank == 0) {          // whole warp, warp-uniform values; one elected lane issues (see the TMA producer)
            constexpr uint32_t idesc = f16_idesc(PAIR ? 2 * BM : BM, BN);
            int it = 0, acc_it = 0;                     // k-block and accumulator-stage counters (across tiles)
            for (int item = first_tile; item < p.num_items; item += tile_step) {
                const Item wi = work_item(p, item);
                // The tensor core truncates (rounds toward zero) when it adds to the fp32 accumulator, a bias that grows
                // with the length of the chain.  Split mode therefore accumulates at most `chunk` k-blocks per TMEM
                // stage and lets the epilogue warps add the chunks in registers with round-to-nearest.
                for (int kb0 = wi.kb_lo; kb0 < wi.kb_hi; kb0 += p.chunk, ++acc_it) {
                    const int a = acc_it & 1;
                    if (PAIR) mbar_wait_cluster(acce_bar(a), ((acc_it >> 1) & 1) ^ 1);
                    else mbar_wait(acce_bar(a), ((acc_it >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(a * BN);
                    const int kb1 = min(kb0 + p.chunk, wi.kb_hi);
                    for (int kb = kb0; kb < kb1; ++kb, ++it) {
                        const int s = it % G::STAGES;
                        mbar_wait(full_bar(s), (it / G::STAGES) & 1);
                        if (PAIR) mbar_wait_cluster(peer_full_bar(s), (it / G::STAGES) & 1);
                        tc_fence_after();
                        const uint32_t a_hi = smem_base + s * G::STAGE;
                        const uint32_t b_hi = a_hi + G::PLANES * A_TILE;
                        const bool main_block = kb < p.kb_main;     // residual blocks multiply by (identity * f): the W_lo term is zero
                        if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            if (p.dbg_skip & 4) break;
                            const uint64_t da = sw128_desc(a_hi + k * 32), db = sw128_desc(b_hi + k * 32);
                            const uint32_t accum = (kb > kb0) || (k != 0);
                            if (PAIR) tc_mma_f16_pair(d_tmem, da, db, idesc, accum);
                            else tc_mma_f16(d_tmem, da, db, idesc, accum);
                            if (SPLIT) {
                                const uint64_t dal = sw128_desc(a_hi + A_TILE + k * 32);
                                if (PAIR) tc_mma_f16_pair(d_tmem, dal, db, idesc, 1);
                                else tc_mma_f16(d_tmem, dal, db, idesc, 1);
                                if (main_block) {
                                    const uint64_t dbl = sw128_desc(b_hi + G::B_LOCAL + k * 32);
                                    if (PAIR) tc_mma_f16_pair(d_tmem, da, dbl, idesc, 1);
                                    else tc_mma_f16(d_tmem, da, dbl, idesc, 1);
                                }
                            }
                        }
                        if (QUAD) tc_commit_mc(empty_bar(s), 0xF);     // this pair is done with stage s: tell all four CTAs (their refills land in both pairs)
                        else if (PAIR) tc_commit_pair(empty_bar(s));   // smem slot reusable in BOTH CTAs once these MMAs retire
                        else tc_commit(empty_bar(s));
                        if (kb == kb1 - 1) {
                            if (QUAD) tc_commit_mc(accf_bar(a), (uint16_t)(3u << leader_rank));   // chunk complete -> the epilogues of this pair's two CTAs
                            else if (PAIR) tc_commit_pair(accf_bar(a));        // chunk complete -> both CTAs' epilogues
                            else tc_commit(accf_bar(a));
                        }
                        }
                        __syncwarp();
                    }
                }
            }
        } else if (PAIR && lane == 0) {
            // peer CTA: tell the leader's MMA thread when this CTA's half of a stage has landed
            int it = 0;
            for (int item = first_tile; item < p.num_items; item += tile_step) {
                const Item wi = work_item(p, item);
                for (int kb = wi.kb_lo; kb < wi.kb_hi; ++kb, ++it) {
                    const int s = it % G::STAGES;
                    mbar_wait(full_bar(s), (it / G::STAGES) & 1);
                    mbar_arrive_remote(map_to_cta(peer_full_bar(s), leader_rank));
                }
            }
        }
    } else {
        // =========================== epilogue (warps 0-7) ===========================
        constexpr int SLAB = G::SLAB;
        constexpr int NSLAB = BN / SLAB;
        constexpr int RB = SLAB * 2;                  // bytes per staged row per plane: 32 / 64 / 128 = the TMA swizzle span
        constexpr uint32_t SW_MASK = RB / 16 - 1;
        const int group = warp >> 2, quarter = warp & 3;      // TMEM lanes [32*quarter, +32) are this warp's
        const bool leader = (threadIdx.x & 127) == 0;         // issues this group's TMA stores
        const int row = quarter * 32 + lane;
        const uint32_t buf0 = staging + (uint32_t)group * G::GROUP_STAGING;     // NBUF staging slabs (hi | lo) per group
        constexpr int MINE = (NSLAB + 1) / 2;         // slabs per group
        int acc_it = 0;
        uint32_t slab_it = 0;                         // slabs this group has staged so far (selects the ping-pong buffer)
        uint32_t tile_it = 0;
        const uint32_t acce_remote0 = PAIR ? map_to_cta(acce_bar(0), leader_rank) : 0u;     // the pair leader's acc_empty barriers, as cluster addresses
        for (int item = first_tile; item < p.num_items; item += tile_step, ++tile_it) {
            const Item wi = work_item(p, item);
            const TileCoord tc = tile_coord<GATHER, PAIR, QUAD>(p, wi.tile, cta_rank, pair_id);
            // This n-tile's scale / shift go through shared memory: every lane needs the same 2*BN values, so the epilogue reads
            // them as broadcast LDS instead of four LDG.128 per 8 channels whose latency sat in front of every FFMA (long-scoreboard
            // stalls were ~20 % of the epilogue warps' samples).  Double buffered by tile parity; the barrier below orders tile i's
            // readers before tile i+2's writers.
            const uint32_t ss = ss_base + (tile_it & 1u) * (uint32_t)(2 * BN * 4);
            for (int e = threadIdx.x; e < 2 * BN; e += EPI_THREADS) {
                const int c = tc.nt * BN + (e < BN ? e : e - BN);
                const float v = c < p.cout_pad ? __ldg((e < BN ? p.scale : p.shift) + c) : 0.f;
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(ss + 4u * e), "f"(v) : "memory");
            }
            bool valid;
            int n, oy, ox;
            if (GATHER) {
                const int m = tc.mt * BM + row;
                valid = m < p.M;
                const int qx = m % p.Qw, t2 = m / p.Qw;
                n = t2 / p.Qh, oy = t2 % p.Qh, ox = qx;
            } else {
                const int qx = tc.qx0 + (row & (p.bw - 1)), qy = tc.qy0 + ((row >> p.lg_bw) & (p.bh - 1));
                n = tc.n0 + (row >> (p.lg_bw + p.lg_bh));
                valid = qx < p.Qw && qy < p.Qh && n < p.N;
                oy = qy * p.out_mul + p.ph_oy[tc.phase];
                ox = qx * p.out_mul + p.ph_ox[tc.phase];
            }
            // ---- drain the accumulator chunks of this tile into registers (fp32, round-to-nearest adds) ----
            uint32_t acc[MINE * SLAB];
            for (int kb0 = wi.kb_lo; kb0 < wi.kb_hi; kb0 += p.chunk, ++acc_it) {
                const int a = acc_it & 1;
                mbar_wait(accf_bar(a), (acc_it >> 1) & 1);
                tc_fence_after();
                const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(a * BN);
                if (kb0 == wi.kb_lo) {
#pragma unroll
                    for (int j = 0; j < MINE; ++j) {
                        const int cc = group + 2 * j;
                        if (cc < NSLAB) {
                            if constexpr (SLAB == 32) {
                                TMEM_LD_32(t_row + cc * SLAB, (acc + j * SLAB));
                            } else {
                                TMEM_LD_16(t_row + cc * SLAB, (acc + j * SLAB));
                            }
                        }
                    }
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                } else {
#pragma unroll
                    for (int j = 0; j < MINE; ++j) {
                        const int cc = group + 2 * j;
                        if (cc < NSLAB) {
                            uint32_t v[SLAB];
                            if constexpr (SLAB == 32) {
                                TMEM_LD_32(t_row + cc * SLAB, v);
                            } else {
                                TMEM_LD_16(t_row + cc * SLAB, v);
                            }
                            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                            for (int i = 0; i < SLAB; ++i) acc[j * SLAB + i] = __float_as_uint(__uint_as_float(acc[j * SLAB + i]) + __uint_as_float(v[i]));
                        }
                    }
                }
                tc_fence_before();
                if (PAIR) mbar_arrive_remote(acce_remote0 + 8u * a);   // leader's barrier counts the epilogue threads of both CTAs
                else mbar_arrive(acce_bar(a));        // this thread is done with the TMEM stage
            }
            named_bar_sync(3, EPI_THREADS);           // scale/shift of this tile are staged (all eight epilogue warps)
            if (p.splits > 1) {
                // ---- split-K: park this item's partial sums; the CTA whose part arrives last adds all parts (in split order, so the
                // result does not depend on who is last) and finalizes the tile ----
                const size_t cta_tile = (size_t)wi.tile * CL + crank;
                // workspace layout per (cta tile, part): [BN / 4 column quads][128 rows] float4 -- a thread's 16-byte accesses, lanes =
                // consecutive rows: every warp access is 512 contiguous bytes
                uint4* wsp = reinterpret_cast<uint4*>(p.sk_ws + (cta_tile * p.splits + wi.sp) * (size_t)(BN * BM)) + row;
                if (!(p.dbg_skip & 16)) {
#pragma unroll
                    for (int j = 0; j < MINE; ++j) {
                        const int cc = group + 2 * j;
                        if (cc < NSLAB) {
#pragma unroll
                            for (int i = 0; i < SLAB / 4; ++i)
                                wsp[(size_t)(cc * (SLAB / 4) + i) * BM] = make_uint4(acc[j * SLAB + 4 * i], acc[j * SLAB + 4 * i + 1], acc[j * SLAB + 4 * i + 2], acc[j * SLAB + 4 * i + 3]);
                        }
                    }
                }
                // release / acquire through ONE thread (the cooperative-groups grid-barrier pattern): the named barrier orders every
                // epilogue thread's partial-sum stores before thread 0's fence + counter increment; on the way back thread 0's fence
                // after the increment orders the other parts' stores before the loads the whole CTA issues behind the second barrier.
                // (256 threads each executing MEMBAR.SC.GPU with their stores in flight cost several microseconds per work item.)
                if (p.dbg_skip & 32) __threadfence();
                named_bar_sync(3, EPI_THREADS);
                if (threadIdx.x == 0) {
                    __threadfence();
                    const unsigned old = atomicAdd(p.sk_cnt + cta_tile, 1u);
                    const unsigned last = old == (unsigned)(p.splits - 1);
                    if (last) p.sk_cnt[cta_tile] = 0u;                          // every part has arrived: leave the slot clean for the next launch
                    __threadfence();
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(ss_base + (uint32_t)G::SS_BYTES), "r"(last) : "memory");
                }
                named_bar_sync(3, EPI_THREADS);
                uint32_t last;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(last) : "r"(ss_base + (uint32_t)G::SS_BYTES));
                if (!last) continue;                                            // warp-uniform (the whole CTA takes the same way)
                if (p.dbg_skip & 32) __threadfence();
                // all of a part's loads are issued before the first add (MINE * SLAB / 4 independent 16-byte loads in flight per thread):
                // the reduction costs about one L2 round trip per part, not one per element
                const uint4* w0 = reinterpret_cast<const uint4*>(p.sk_ws + cta_tile * p.splits * (size_t)(BN * BM)) + row;
                for (int sp = 0; sp < ((p.dbg_skip & 128) ? 1 : p.splits); ++sp) {
                    const uint4* q = w0 + (size_t)sp * (BN * BM / 4);
                    uint4 v[MINE * (SLAB / 4)];
#pragma unroll
                    for (int j = 0; j < MINE; ++j) {
                        const int cc = group + 2 * j;
                        if (cc < NSLAB) {
#pragma unroll
                            for (int i = 0; i < SLAB / 4; ++i) v[j * (SLAB / 4) + i] = __ldcg(q + (size_t)(cc * (SLAB / 4) + i) * BM);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < MINE; ++j) {
                        const int cc = group + 2 * j;
                        if (cc < NSLAB) {
#pragma unroll
                            for (int i = 0; i < SLAB / 4; ++i) {
                                const uint4 t = v[j * (SLAB / 4) + i];
                                uint32_t* a4 = acc + j * SLAB + 4 * i;
                                if (sp == 0) {
                                    a4[0] = t.x, a4[1] = t.y, a4[2] = t.z, a4[3] = t.w;
                                } else {
                                    a4[0] = __float_as_uint(__uint_as_float(a4[0]) + __uint_as_float(t.x));
                                    a4[1] = __float_as_uint(__uint_as_float(a4[1]) + __uint_as_float(t.y));
                                    a4[2] = __float_as_uint(__uint_as_float(a4[2]) + __uint_as_float(t.z));
                                    a4[3] = __float_as_uint(__uint_as_float(a4[3]) + __uint_as_float(t.w));
                                }
                            }
                        }
                    }
                }
            }
            // ---- finalize: scale/shift, activation, split, store ----
#pragma unroll
            for (int j = 0; j < MINE; ++j) {
                const int cc = group + 2 * j;
                const int cbase = tc.nt * BN + cc * SLAB;
                if (cc >= NSLAB || cbase >= p.cout_p8) continue;    // no such slab / whole slab is channel padding (uniform across the group)
                float o[SLAB];
#pragma unroll
                for (int g8 = 0; g8 < SLAB / 8; ++g8) {
                    const int c0 = cbase + g8 * 8;
                    if (c0 < p.cout_p8) {
                        float scv[8], shv[8];
                        const uint32_t so = ss + 4u * (uint32_t)(cc * SLAB + g8 * 8);
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(scv[0]), "=f"(scv[1]), "=f"(scv[2]), "=f"(scv[3]) : "r"(so));
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(scv[4]), "=f"(scv[5]), "=f"(scv[6]), "=f"(scv[7]) : "r"(so + 16u));
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(shv[0]), "=f"(shv[1]), "=f"(shv[2]), "=f"(shv[3]) : "r"(so + 4u * BN));
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(shv[4]), "=f"(shv[5]), "=f"(shv[6]), "=f"(shv[7]) : "r"(so + 4u * BN + 16u));
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            // branch-free activation: max(v, min(s*v, cap)) with s = 1 (none), 0 (ReLU, cap 0: inf * 0 = NaN must not leak), slope (LeakyReLU)
                            const float v = fmaf(__uint_as_float(acc[j * SLAB + g8 * 8 + k]), scv[k], shv[k]);
                            o[g8 * 8 + k] = fmaxf(v, fminf(v * p.act_mul, p.act_cap));
                        }
                        if (p.out_f32 && valid) {
#pragma unroll
                            for (int k = 0; k < 8; ++k)
                                if (c0 + k < p.cout) p.out_f32[(((size_t)n * p.cout + c0 + k) * p.Ho + oy) * p.Wo + ox] = o[g8 * 8 + k];
                        }
                    }
                }
                if (p.out_hi) {
                    // the TMA store that last read this staging slab must have drained it (NBUF = 2: the store before the last one)
                    const uint32_t buf = buf0 + (NBUF > 1 ? (slab_it & 1u) * (2 * G::SLAB_PLANE) : 0u);
                    ++slab_it;
                    if (leader) bulk_wait_read<NBUF - 1>();
                    named_bar_sync(1 + group, 128);
#pragma unroll
                    for (int g8 = 0; g8 < SLAB / 8; ++g8) {
                        if (cbase + g8 * 8 >= p.cout_p8) break;
                        uint32_t hv[4], lv[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) split_f32x2(o[g8 * 8 + 2 * k], o[g8 * 8 + 2 * k + 1], hv[k], lv[k]);
                        uint32_t off = (uint32_t)row * RB + (uint32_t)g8 * 16u;
                        off ^= ((off >> 7) & SW_MASK) << 4;
                        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(buf + off), "r"(hv[0]), "r"(hv[1]), "r"(hv[2]), "r"(hv[3]) : "memory");
                        if (p.out_lo)
                            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(buf + G::SLAB_PLANE + off), "r"(lv[0]), "r"(lv[1]), "r"(lv[2]), "r"(lv[3]) : "memory");
                    }
                    fence_proxy_async();              // generic-proxy smem writes -> visible to the TMA (async proxy)
                    named_bar_sync(1 + group, 128);
                    if (leader && !(p.dbg_skip & 8)) {
                        tma_store_4d(&p.out_map[tc.phase][0], buf, cbase, tc.qx0, tc.qy0, tc.n0);
                        if (p.out_lo) tma_store_4d(&p.out_map[tc.phase][1], buf + G::SLAB_PLANE, cbase, tc.qx0, tc.qy0, tc.n0);
                        bulk_commit();
                    }
                }
            }
        }
        if (leader) bulk_wait_all();                  // every output byte has left shared memory and is globally visible
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();          // neither CTA may leave (or free TMEM) while its peer can still touch its smem / barriers
    if (warp == MMA_WARP) {
        tc_fence_after();
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)G::TMEM_COLS) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)G::TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// SIMT cross-check kernel: same inputs, same outputs, fp32 FMAs.
// ------------------------------------------------------------------------------------------------
__global__ void conv_simt_kernel(const __grid_constant__ ConvK p) {
    const size_t total = (size_t)p.n_phase * p.M * p.cout_p8;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int co = (int)(idx % p.cout_p8);
        const size_t r = idx / p.cout_p8;
        const int m = (int)(r % p.M), ph = (int)(r / p.M);
        const int qx = m % p.Qw, t2 = m / p.Qw;
        const int qy = t2 % p.Qh, n = t2 / p.Qh;
        float acc = 0.f;
        const float* wp = p.w_gemm + (size_t)ph * p.k_pad * p.cout_pad + co;
        for (int t = 0; t < p.taps; ++t) {
            const int iy = qy * p.in_stride + p.tap_dy[ph][t], ix = qx * p.in_stride + p.tap_dx[ph][t];
            if (iy < 0 || iy >= p.H || ix < 0 || ix >= p.W) continue;
            const size_t off = ((size_t)(n * p.H + iy) * p.Wp + ix + p.xoff) * p.in_ld + p.in_coff;
            const float* wt = wp + ((size_t)t * p.cin_k + (size_t)(t / p.kw) * p.row_pad) * p.cout_pad;
            for (int ci = 0; ci < p.cin_p; ++ci) {
                float a = __half2float(p.in_hi[off + ci]);
                if (p.in_lo) a += __half2float(p.in_lo[off + ci]);
                acc = fmaf(a, __ldg(wt + (size_t)ci * p.cout_pad), acc);
            }
        }
        const int oy = qy * p.out_mul + p.ph_oy[ph], ox = qx * p.out_mul + p.ph_ox[ph];
        const size_t opix = ((size_t)n * p.Ho + oy) * p.Wo + ox;
        float o = fmaf(acc, p.scale[co], p.shift[co]);
        if (p.res_hi) {
            o += __half2float(p.res_hi[opix * p.res_ld + p.res_coff + co]);
            if (p.res_lo) o += __half2float(p.res_lo[opix * p.res_ld + p.res_coff + co]);
        }
        o = apply_act(o, p.act, p.slope);
        if (p.out_hi) {
            __half h, l;
            split_f32(o, h, l);
            p.out_hi[opix * p.out_ld + p.out_coff + co] = h;
            if (p.out_lo) p.out_lo[opix * p.out_ld + p.out_coff + co] = l;
        }
        if (p.out_f32 && co < p.cout) p.out_f32[(((size_t)n * p.cout + co) * p.Ho + oy) * p.Wo + ox] = o;
    }
}

template <int BN, bool SPLIT, bool GATHER, bool PAIR, int NBUF = 1, int OCC = 1, bool QUAD = false>
static int launch_tc(const ConvK& k, cudaStream_t s) {
    using G = Geo<BN, SPLIT, PAIR, NBUF, OCC>;
    static bool attr_set[MAX_DEVICES];   // per instantiation and device
    if (ensure_dyn_smem(conv_tc_kernel<BN, SPLIT, GATHER, PAIR, NBUF, OCC, QUAD>, G::SMEM, attr_set, "conv_tc_kernel")) return 1;
    int sms = device_sms() * OCC;        // OCC = 2: two resident CTAs per SM
    if (const int cap = dev_knob(FT_DBG_MAX_CTAS)) sms = cap < sms ? (cap < 2 ? 2 : cap) : sms;      // dev experiment: persistent grid on part of the GPU
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (!dev_knob(FT_DBG_NO_PDL)) {      // programmatic dependent launch: see pdl_wait() in the kernel
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (PAIR) {
        // one cluster of two CTAs (the two SMs of a TPC) -- QUAD: four CTAs, two TPCs -- per work item; k.num_tiles already counts cluster tiles
        constexpr int CL = QUAD ? 4 : 2;
        const int clusters = k.num_items < sms / CL ? k.num_items : sms / CL;
        cfg.gridDim = dim3(CL * clusters), cfg.blockDim = dim3(THREADS_TMA);
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = CL, attr[na].val.clusterDim.y = 1, attr[na].val.clusterDim.z = 1;
        ++na;
    } else {
        cfg.gridDim = dim3(k.num_items < sms ? k.num_items : sms), cfg.blockDim = dim3(GATHER ? THREADS_GATHER : THREADS_TMA);
    }
    cfg.dynamicSmemBytes = G::SMEM, cfg.stream = s, cfg.attrs = attr, cfg.numAttrs = na;
    cudaError_t e = cudaLaunchKernelEx(&cfg, conv_tc_kernel<BN, SPLIT, GATHER, PAIR, NBUF, OCC, QUAD>, k);
    if (e != cudaSuccess) {
        set_error("conv_tc_kernel<%d,%d,%d,%d>: launch failed: %s", BN, (int)SPLIT, (int)GATHER, (int)PAIR, cudaGetErrorString(e));
        return 1;
    }
    return check_launch(PAIR ? "conv_tc_kernel(pair)" : "conv_tc_kernel");
}

static int check_view(const ft_act& a, const char* name) {
    FT_REQUIRE(a.c % 8 == 0 && a.ld % 8 == 0 && a.coff % 8 == 0 && a.c > 0, "%s: c/ld/coff must be positive multiples of 8 (c=%d ld=%d coff=%d)", name, a.c, a.ld, a.coff);
    FT_REQUIRE(a.coff + a.c <= a.ld, "%s: view exceeds pixel stride", name);
    FT_REQUIRE(((uintptr_t)a.hi & 15) == 0 && ((uintptr_t)a.lo & 15) == 0, "%s: planes must be 16-byte aligned", name);
    FT_REQUIRE(a.n > 0 && a.h > 0 && a.w > 0, "%s: empty tensor", name);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// TMA tensor maps (driver entry point resolved through the runtime: no link-time libcuda dependency)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

// fp16 tensor [d3][d2][d1][d0] with d0 contiguous; `stride` in ELEMENTS for d1..d3; box in elements.
int make_map(CUtensorMap* m, const void* base, const long long dim[4], const long long stride[3], const int box[4], int swizzle_bytes, const char* what) {
    EncodeTiledFn enc = encode_fn();
    FT_REQUIRE(enc, "conv2d: cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t gd[4], gs[3];
    cuuint32_t bx[4], es[4] = {1, 1, 1, 1};
    for (int i = 0; i < 4; ++i) gd[i] = (cuuint64_t)dim[i], bx[i] = (cuuint32_t)box[i];
    for (int i = 0; i < 3; ++i) gs[i] = (cuuint64_t)stride[i] * 2;
    const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FT_REQUIRE(r == CUDA_SUCCESS, "conv2d: cuTensorMapEncodeTiled(%s) failed with %d (dims %lld,%lld,%lld,%lld strides %lld,%lld,%lld box %d,%d,%d,%d)", what, (int)r,
               dim[0], dim[1], dim[2], dim[3], stride[0], stride[1], stride[2], box[0], box[1], box[2], box[3]);
    return 0;
}

// Pixel box of a tile: bw*bh*bn = 128, powers of two, fewest tiles (least padding); ties prefer long rows.
static void choose_box(int Qw, int Qh, int N, int& bw, int& bh, int& bn) {
    long long best = -1;
    for (int lw = 7; lw >= 0; --lw)
        for (int lh = 7 - lw; lh >= 0; --lh) {
            const int w = 1 << lw, h = 1 << lh, n = 128 / (w * h);
            const long long tiles = (long long)ceil_div(Qw, w) * ceil_div(Qh, h) * ceil_div(N, n);
            if (best < 0 || tiles < best) best = tiles, bw = w, bh = h, bn = n;
        }
}

static int ilog2(int v) {
    int l = 0;
    while ((1 << l) < v) ++l;
    return l;
}

}  // namespace ft

using namespace ft;

extern "C" int ft_conv2d_fprop(const ft_conv_desc* d, void* stream) {
    FT_REQUIRE(d, "conv2d: null descriptor");
    FT_REQUIRE(d->in.hi, "conv2d: null input");
    if (int e = check_view(d->in, "conv2d.in")) return e;
    FT_REQUIRE(d->out.hi || d->out_f32, "conv2d: no output requested");
    FT_REQUIRE(d->scale && d->shift, "conv2d: scale/shift are required");
    FT_REQUIRE(d->cout > 0 && d->cout_pad >= d->cout && d->cout_pad % 8 == 0, "conv2d: bad cout/cout_pad (%d/%d)", d->cout, d->cout_pad);
    FT_REQUIRE(d->precision == FT_PREC_SPLIT || d->precision == FT_PREC_FAST, "conv2d: bad precision %d", d->precision);
    FT_REQUIRE(d->precision == FT_PREC_FAST || d->in.lo, "conv2d: split precision needs the input lo plane");
    FT_REQUIRE(d->k_mode == FT_K_PACKED || d->k_mode == FT_K_TAP64 || d->k_mode == FT_K_ROW, "conv2d: bad k_mode %d", d->k_mode);
    const bool split = d->precision == FT_PREC_SPLIT;

    ConvK k = {};
    k.in_hi = (const __half*)d->in.hi;
    k.in_lo = split ? (const __half*)d->in.lo : nullptr;
    FT_REQUIRE(d->in_xpad_l >= 0 && d->in_xpad_r >= 0 && d->in.w > d->in_xpad_l + d->in_xpad_r, "conv2d: bad input x-padding %d/%d for width %d", d->in_xpad_l, d->in_xpad_r, d->in.w);
    k.N = d->in.n, k.H = d->in.h, k.Wp = d->in.w, k.xoff = d->in_xpad_l, k.W = d->in.w - d->in_xpad_l - d->in_xpad_r;
    k.in_ld = d->in.ld, k.in_coff = d->in.coff, k.cin_p = d->in.c, k.cpt = d->in.c / 8;
    k.kw = d->kind == FT_CONV ? d->kw : 2, k.pad = d->pad;
    if (d->kind == FT_CONV) {
        FT_REQUIRE(d->kh > 0 && d->kw > 0 && d->kh * d->kw <= 64 && d->stride > 0 && d->pad >= 0, "conv2d: unsupported window %dx%d s%d p%d", d->kh, d->kw, d->stride, d->pad);
        k.n_phase = 1, k.taps = d->kh * d->kw;
        const int pad_hi = d->pad_hi >= 0 ? d->pad_hi : d->pad;
        FT_REQUIRE(pad_hi <= d->pad, "conv2d: pad_hi %d must not exceed pad %d", pad_hi, d->pad);
        k.Ho = (k.H + d->pad + pad_hi - d->kh) / d->stride + 1;
        k.Wo = (k.W + d->pad + pad_hi - d->kw) / d->stride + 1;
        k.Qh = k.Ho, k.Qw = k.Wo, k.in_stride = d->stride, k.out_mul = 1;
        for (int r = 0; r < d->kh; ++r)
            for (int s = 0; s < d->kw; ++s) {
                k.tap_dy[0][r * d->kw + s] = (int8_t)(r - d->pad);
                k.tap_dx[0][r * d->kw + s] = (int8_t)(s - d->pad);
            }
    } else if (d->kind == FT_DECONV_K4S2P1) {
        // out[oy] += in[iy]*w[ky], oy = 2*iy - 1 + ky.  Output parity py sees ky = 1 - py + 2a (a = 0,1) at iy = qy + py - a.
        k.n_phase = 4, k.taps = 4;
        k.Ho = 2 * k.H, k.Wo = 2 * k.W, k.Qh = k.H, k.Qw = k.W, k.in_stride = 1, k.out_mul = 2;
        for (int py = 0; py < 2; ++py)
            for (int px = 0; px < 2; ++px) {
                const int ph = py * 2 + px;
                k.ph_oy[ph] = (int8_t)py, k.ph_ox[ph] = (int8_t)px;
                for (int a = 0; a < 2; ++a)
                    for (int b = 0; b < 2; ++b) {
                        k.tap_dy[ph][a * 2 + b] = (int8_t)(py - a);
                        k.tap_dx[ph][a * 2 + b] = (int8_t)(px - b);
                    }
            }
    } else {
        FT_REQUIRE(false, "conv2d: unknown kind %d", d->kind);
    }
    FT_REQUIRE(k.Ho > 0 && k.Wo > 0, "conv2d: empty output");
    // K layout: FT_K_PACKED k = tap*cin_p + ci ; FT_K_TAP64 k = tap*cin_k + ci with cin_k = cin_p rounded up to 64 ;
    //           FT_K_ROW    k = r*64 + s*cin_p + ci (kernel row r, column s)
    k.cin_k = d->k_mode == FT_K_TAP64 ? round_up(k.cin_p, BK) : k.cin_p;
    k.cb = k.cin_k / BK;
    k.k_pad = d->k_pad;
    if (d->k_mode == FT_K_ROW) {
        FT_REQUIRE(d->kind == FT_CONV && d->kw * k.cin_p <= BK && (d->stride == 1 || d->stride == 2), "conv2d: FT_K_ROW needs a plain convolution with kw*Cin <= 64 and stride 1 or 2");
        FT_REQUIRE(d->in.ld == d->in.c && d->in.coff == 0, "conv2d: FT_K_ROW needs a dense input view (ld == c)");
        FT_REQUIRE(d->in_xpad_l >= d->pad && d->in_xpad_r * k.cin_p >= d->pad * k.cin_p + BK - d->kw * k.cin_p,
                   "conv2d: FT_K_ROW needs zero columns in the input rows: left >= pad, right >= pad + (64 - kw*Cin)/Cin (got %d/%d)", d->in_xpad_l, d->in_xpad_r);
        FT_REQUIRE(!d->res.hi, "conv2d: FT_K_ROW does not take a residual");
        k.row_mode = 1, k.row_pad = BK - d->kw * k.cin_p, k.cb = 1;
        FT_REQUIRE(k.k_pad == d->kh * BK, "conv2d: k_pad %d does not match kh*64 = %d (FT_K_ROW)", k.k_pad, d->kh * BK);
    } else {
        FT_REQUIRE(k.k_pad == round_up(k.taps * k.cin_k, BK), "conv2d: k_pad %d does not match taps*cin_k = %d (k_mode %d)", k.k_pad, k.taps * k.cin_k, d->k_mode);
    }
    k.kb_main = k.k_pad / BK;
    k.cout = d->cout, k.cout_p8 = round_up(d->cout, 8), k.cout_pad = d->cout_pad;
    k.scale = d->scale, k.shift = d->shift, k.act = d->act, k.slope = d->slope;
    FT_REQUIRE(d->act != FT_ACT_LEAKY || (d->slope >= 0.f && d->slope <= 1.f), "conv2d: LeakyReLU slope must lie in [0, 1] (got %g)", (double)d->slope);
    k.act_mul = d->act == FT_ACT_RELU ? 0.f : d->act == FT_ACT_LEAKY ? d->slope : 1.f;
    k.act_cap = d->act == FT_ACT_RELU ? 0.f : INFINITY;
    k.w_tile = (const uint8_t*)d->w_tile, k.w_gemm = d->w_gemm;
    if (d->out.hi) {
        if (int e = check_view(d->out, "conv2d.out")) return e;
        FT_REQUIRE(d->out.n == k.N && d->out.h == k.Ho && d->out.w == k.Wo, "conv2d: output view is %dx%dx%d, expected %dx%dx%d", d->out.n, d->out.h, d->out.w, k.N, k.Ho, k.Wo);
        FT_REQUIRE(d->out.c >= k.cout_p8, "conv2d: output view has %d channels, need %d", d->out.c, k.cout_p8);
        k.out_hi = (__half*)d->out.hi, k.out_lo = (__half*)d->out.lo, k.out_ld = d->out.ld, k.out_coff = d->out.coff;
    }
    k.out_f32 = d->out_f32;
    if (d->res.hi) {
        if (int e = check_view(d->res, "conv2d.res")) return e;
        FT_REQUIRE(d->res.n == k.N && d->res.h == k.Ho && d->res.w == k.Wo && d->res.c >= k.cout_p8, "conv2d: residual shape mismatch");
        k.res_hi = (const __half*)d->res.hi, k.res_lo = split ? (const __half*)d->res.lo : nullptr, k.res_ld = d->res.ld, k.res_coff = d->res.coff;
        FT_REQUIRE(!split || d->res.lo, "conv2d: split precision needs the residual lo plane");
    }
    const long long M = (long long)k.N * k.Qh * k.Qw;
    FT_REQUIRE(M * 4 < (1ll << 31) && (long long)k.N * k.H * k.W < (1ll << 31), "conv2d: tensor too large for 32-bit pixel indices");
    k.M = (int)M;
    cudaStream_t s = (cudaStream_t)stream;

    if (d->impl == FT_IMPL_SIMT) {
        FT_REQUIRE(d->w_gemm, "conv2d: SIMT implementation needs w_gemm");
        const size_t total = (size_t)k.n_phase * k.M * k.cout_p8;
        size_t g = (total + 255) / 256;
        if (g > (size_t)device_sms() * 32) g = (size_t)device_sms() * 32;
        conv_simt_kernel<<<(int)g, 256, 0, s>>>(k);
        return check_launch("conv_simt_kernel");
    }
    FT_REQUIRE(d->impl == FT_IMPL_TCGEN05, "conv2d: unknown impl %d", d->impl);
    FT_REQUIRE(d->w_tile && ((uintptr_t)d->w_tile & 15) == 0, "conv2d: tcgen05 implementation needs a 16-byte aligned w_tile");
    const int BN = d->block_n;
    FT_REQUIRE(BN == 16 || BN == 32 || BN == 64 || BN == 128 || BN == 256, "conv2d: block_n must be 16, 32, 64, 128 or 256 (got %d)", BN);
    FT_REQUIRE(d->cout_pad % BN == 0, "conv2d: cout_pad %d is not a multiple of block_n %d", d->cout_pad, BN);
    k.n_tiles = d->cout_pad / BN;
    k.n_tiles_w = k.n_tiles;
    const bool gather = d->k_mode == FT_K_PACKED;
    const bool quad = d->cta_pair == 2;
    const int res_blocks = d->res.hi ? ceil_div(BN, BK) : 0;
    FT_REQUIRE(d->res_blocks == res_blocks, "conv2d: the weight image carries %d residual k-blocks per tile, this call needs %d (residual %s, block_n %d)",
               d->res_blocks, res_blocks, d->res.hi ? "given" : "absent", BN);
    k.kb_total = k.kb_main + res_blocks;
    k.dbg_skip = dev_knob(FT_DBG_SKIP);                                      // ft_set_debug(): 0 unless a tools/dev experiment set it
    const int acc_chunk = dev_knob(FT_DBG_CHUNK) > 0 ? dev_knob(FT_DBG_CHUNK) : ACC_CHUNK;
    k.chunk = (split && k.kb_total > acc_chunk) ? acc_chunk : k.kb_total;   // k-blocks accumulated per TMEM stage (see the MMA warp)
    const int slab = (BN >= 64 && d->ctas_per_sm != 2) ? 32 : 16;      // channels per epilogue staging slab = inner box of the output maps (Geo::SLAB)

    if (gather) {
        FT_REQUIRE(d->kind == FT_CONV && !d->res.hi, "conv2d: FT_K_PACKED (gather variant) supports plain convolutions without residual only");
        k.m_tiles = ceil_div(k.M, BM);
        if (k.out_hi) {   // output as a [M, C] matrix: a tile is 128 consecutive pixels
            const long long dim[4] = {k.cout_p8, k.M, 1, 1}, str[3] = {k.out_ld, (long long)k.out_ld * k.M, (long long)k.out_ld * k.M};
            const int box[4] = {slab, BM, 1, 1};
            if (int e = make_map(&k.out_map[0][0], k.out_hi + k.out_coff, dim, str, box, slab * 2, "out.hi")) return e;
            if (k.out_lo)
                if (int e = make_map(&k.out_map[0][1], k.out_lo + k.out_coff, dim, str, box, slab * 2, "out.lo")) return e;
        }
    } else {
        FT_REQUIRE(k.in_stride == 1 || k.in_stride == 2, "conv2d: the TMA variant supports stride 1 and 2 (got %d)", k.in_stride);
        choose_box(k.Qw, k.Qh, k.N, k.bw, k.bh, k.bn);
        k.lg_bw = ilog2(k.bw), k.lg_bh = ilog2(k.bh);
        k.tiles_w = ceil_div(k.Qw, k.bw), k.tiles_h = ceil_div(k.Qh, k.bh);
        k.m_tiles = k.tiles_w * k.tiles_h * ceil_div(k.N, k.bn);
        const int abox[4] = {BK, k.bw, k.bh, k.bn};
        const int st = k.in_stride;
        if (k.row_mode) {
            // rows of the A tile = 64 contiguous fp16 of the x-padded input starting at pixel (s*qy + r - pad, s*qx - pad):
            // dim1 walks qx with an (overlapping) stride of s pixels, dim2 walks the input rows of one parity, dim3 images
            for (int py = 0; py < st; ++py) {
                const long long dim[4] = {BK, k.Qw, (k.H - py + st - 1) / st, k.N};
                const long long str[3] = {(long long)st * k.in_ld, (long long)st * k.Wp * k.in_ld, (long long)k.H * k.Wp * k.in_ld};
                const size_t off = (size_t)(py * k.Wp + k.xoff - k.pad) * k.in_ld;
                if (int e = make_map(&k.in_map[py * 2][0], k.in_hi + off, dim, str, abox, 128, "in.hi(row)")) return e;
                if (split)
                    if (int e = make_map(&k.in_map[py * 2][1], k.in_lo + off, dim, str, abox, 128, "in.lo(row)")) return e;
            }
        }
        for (int view = 0; view < st * st && !k.row_mode; ++view) {
            const int py = view / st, px = view % st;
            if (px >= k.W || py >= k.H) continue;
            const long long dim[4] = {k.cin_p, (k.W - px + st - 1) / st, (k.H - py + st - 1) / st, k.N};
            const long long str[3] = {(long long)st * k.in_ld, (long long)st * k.Wp * k.in_ld, (long long)k.H * k.Wp * k.in_ld};
            const size_t off = (size_t)(py * k.Wp + px + k.xoff) * k.in_ld + k.in_coff;
            if (int e = make_map(&k.in_map[view][0], k.in_hi + off, dim, str, abox, 128, "in.hi")) return e;
            if (split)
                if (int e = make_map(&k.in_map[view][1], k.in_lo + off, dim, str, abox, 128, "in.lo")) return e;
            if (quad) {      // the same view with half pixel boxes: rows [0,64) and [64,128) of a tile are the two halves of its slowest box dimension
                int hbox[4] = {BK, k.bw, k.bh, k.bn};
                k.half_dx = k.half_dy = k.half_dn = 0;
                if (k.bn > 1) hbox[3] = k.bn / 2, k.half_dn = k.bn / 2;
                else if (k.bh > 1) hbox[2] = k.bh / 2, k.half_dy = k.bh / 2;
                else hbox[1] = k.bw / 2, k.half_dx = k.bw / 2;
                if (int e = make_map(&k.in_map_h[view][0], k.in_hi + off, dim, str, hbox, 128, "in.hi(half)")) return e;
                if (split)
                    if (int e = make_map(&k.in_map_h[view][1], k.in_lo + off, dim, str, hbox, 128, "in.lo(half)")) return e;
            }
        }
        if (k.out_hi) {
            const int mul = k.out_mul;
            const int obox[4] = {slab, k.bw, k.bh, k.bn};
            for (int ph = 0; ph < k.n_phase; ++ph) {
                const long long dim[4] = {k.cout_p8, k.Qw, k.Qh, k.N};
                const long long str[3] = {(long long)mul * k.out_ld, (long long)mul * k.Wo * k.out_ld, (long long)k.Ho * k.Wo * k.out_ld};
                const size_t off = (size_t)(k.ph_oy[ph] * k.Wo + k.ph_ox[ph]) * k.out_ld + k.out_coff;
                if (int e = make_map(&k.out_map[ph][0], k.out_hi + off, dim, str, obox, slab * 2, "out.hi")) return e;
                if (k.out_lo)
                    if (int e = make_map(&k.out_map[ph][1], k.out_lo + off, dim, str, obox, slab * 2, "out.lo")) return e;
            }
        }
        if (d->res.hi) {
            const long long dim[4] = {k.cout_p8, k.Wo, k.Ho, k.N};
            const long long str[3] = {k.res_ld, (long long)k.Wo * k.res_ld, (long long)k.Ho * k.Wo * k.res_ld};
            if (int e = make_map(&k.res_map[0], k.res_hi + k.res_coff, dim, str, abox, 128, "res.hi")) return e;
            if (split)
                if (int e = make_map(&k.res_map[1], k.res_lo + k.res_coff, dim, str, abox, 128, "res.lo")) return e;
        }
    }
    k.rev = d->tile_order != 0;
    const bool pair = d->cta_pair != 0;
    const bool epi2 = d->epilogue_buffers == 2;
    FT_REQUIRE(d->epilogue_buffers >= 0 && d->epilogue_buffers <= 2 && (!epi2 || (BN >= 64 && BN <= 128 && !gather)),
               "conv2d: epilogue_buffers must be 0/1 or 2 (2: TMA variant with block_n 64 or 128)");
    FT_REQUIRE(!pair || (!gather && BN >= 64), "conv2d: cta_pair needs the TMA variant and block_n >= 64");
    FT_REQUIRE(d->cta_pair >= 0 && d->cta_pair <= 2, "conv2d: cta_pair must be 0, 1 or 2");
    if (quad) {
        FT_REQUIRE(!k.row_mode && (BN == 64 || BN == 128) && k.n_tiles % 2 == 0 && !epi2 && d->ctas_per_sm != 2 && d->splitk <= 1,
                   "conv2d: cta_pair = 2 (two CTA pairs sharing the A tile) needs FT_K_TAP64, block_n 64 or 128, an even number of n-tiles (got %d), "
                   "one staging slab, one CTA per SM and no split-K", k.n_tiles);
        k.n_tiles /= 2;                 // a work item spans two neighbouring n-tiles
    }
    k.num_tiles = k.n_phase * (pair ? (k.m_tiles + 1) / 2 : k.m_tiles) * k.n_tiles;
    // split-K: `splitk` parts of every tile's K range become separate work items (see work_item / the epilogue)
    k.splits = 1, k.kb_split = k.kb_total;
    if (d->splitk > 1) {
        FT_REQUIRE(!gather, "conv2d: split-K is not available for the FT_K_PACKED (gather) variant");
        FT_REQUIRE(d->splitk <= 16 && d->splitk <= k.kb_total, "conv2d: splitk %d exceeds 16 or the %d k-blocks of this layer", d->splitk, k.kb_total);
        k.kb_split = ceil_div(k.kb_total, d->splitk);
        k.splits = ceil_div(k.kb_total, k.kb_split);
        const size_t cta_tiles = (size_t)k.num_tiles * (pair ? 2 : 1);
        FT_REQUIRE(d->splitk_ws && ((uintptr_t)d->splitk_ws & 15) == 0 && d->splitk_counters, "conv2d: split-K needs a 16-byte aligned workspace and a counter array");
        FT_REQUIRE(d->splitk_ws_bytes >= cta_tiles * k.splits * (size_t)BN * BM * 4 && (size_t)d->splitk_counters_len >= cta_tiles,
                   "conv2d: split-K workspace too small (need %zu bytes and %zu counters, got %llu / %d)", cta_tiles * k.splits * (size_t)BN * BM * 4, cta_tiles,
                   (unsigned long long)d->splitk_ws_bytes, d->splitk_counters_len);
        k.sk_ws = (float*)d->splitk_ws, k.sk_cnt = (unsigned*)d->splitk_counters;
    }
    k.num_items = k.num_tiles * k.splits;
    k.fd_num_tiles = make_fastdiv(k.num_tiles > 0 ? k.num_tiles : 1);
    k.fd_n_tiles = make_fastdiv(k.n_tiles), k.fd_m_units = make_fastdiv(pair ? (k.m_tiles + 1) / 2 : k.m_tiles);
    k.fd_tiles_w = make_fastdiv(k.tiles_w > 0 ? k.tiles_w : 1), k.fd_tiles_h = make_fastdiv(k.tiles_h > 0 ? k.tiles_h : 1);

#define FT_CASE(BN_)                                                                                                          \
    case BN_:                                                                                                                 \
        if (gather) return split ? launch_tc<BN_, true, true, false>(k, s) : launch_tc<BN_, false, true, false>(k, s);        \
        return split ? launch_tc<BN_, true, false, false>(k, s) : launch_tc<BN_, false, false, false>(k, s);
#define FT_CASE_PAIR(BN_)                                                                                                     \
    case BN_:                                                                                                                 \
        if (epi2 && !gather) {                                                                                                \
            if (pair) return split ? launch_tc<BN_, true, false, true, 2>(k, s) : launch_tc<BN_, false, false, true, 2>(k, s);   \
            return split ? launch_tc<BN_, true, false, false, 2>(k, s) : launch_tc<BN_, false, false, false, 2>(k, s);        \
        }                                                                                                                     \
        if (pair) return split ? launch_tc<BN_, true, false, true>(k, s) : launch_tc<BN_, false, false, true>(k, s);          \
        if (gather) return split ? launch_tc<BN_, true, true, false>(k, s) : launch_tc<BN_, false, true, false>(k, s);        \
        return split ? launch_tc<BN_, true, false, false>(k, s) : launch_tc<BN_, false, false, false>(k, s);
    if (quad) {
        if (BN == 64) return split ? launch_tc<64, true, false, true, 1, 1, true>(k, s) : launch_tc<64, false, false, true, 1, 1, true>(k, s);
        return split ? launch_tc<128, true, false, true, 1, 1, true>(k, s) : launch_tc<128, false, false, true, 1, 1, true>(k, s);
    }
    if (d->ctas_per_sm == 2) {
        FT_REQUIRE(BN == 64 && pair && !gather && !epi2, "conv2d: ctas_per_sm = 2 is available for block_n 64 on CTA pairs (one staging slab)");
        return split ? launch_tc<64, true, false, true, 1, 2>(k, s) : launch_tc<64, false, false, true, 1, 2>(k, s);
    }
    FT_REQUIRE(d->ctas_per_sm >= 0 && d->ctas_per_sm <= 2, "conv2d: ctas_per_sm must be 0/1 or 2");
    switch (BN) {
        FT_CASE(16)
        FT_CASE(32)
        FT_CASE_PAIR(64)
        FT_CASE_PAIR(128)
        case 256:      // CTA pairs only (a single CTA could not hold two stages of a 256-wide weight tile), no ping-pong staging
            FT_REQUIRE(pair, "conv2d: block_n 256 needs cta_pair");
            return split ? launch_tc<256, true, false, true>(k, s) : launch_tc<256, false, false, true>(k, s);
    }
#undef FT_CASE
#undef FT_CASE_PAIR
    return 0;
}
