// ft_corr.cu -- the FlowNetC cost volume (Correlation: pad 20, kernel 1, max displacement 20, stride1 1,
// stride2 2; reference: correlation_package/src/correlation_cuda_kernel.cu:34-106) as a banded GEMM on the
// tcgen05 tensor cores.  sm_100a only.
//
//   out[n, (tj+10)*21 + (ti+10), y, x] = 1/C * sum_c in1[n,c,y,x] * in2[n,c,y+2tj,x+2ti]     (zero outside the image)
//
// Because stride2 = 2, pixel (y,x) only ever meets pixels of its own row/column parity.  Each of the four parity
// classes (py,px) is therefore an independent stride-1 correlation (displacements |tj|,|ti| <= 10) between the
// sub-images in1' = in1[:, py::2, px::2] and in2' likewise ([H/2, W/2] pixels of C channels).  A work tile is
//   (image n, parity class, 4 rows a0..a0+3 of in1')  =  128 pixels x C channels  -> the GEMM's A operand (M = 128),
// and it needs in2' rows a0-10 .. a0+13: they are streamed as "chunks" of 4 rows x 32 columns = 128 pixels (N = 128):
//   D[128 x 128] = A . B^T  (K = C), the all-pairs dot products of the 128 in1' pixels with the 128 in2' pixels;
//   the 21-wide diagonal band |b2 - b| <= 10 of every (row a, row a+tj) block is the wanted output, the rest is dropped
//   (useful fraction about 2/3 x 21/32; the operands are read ONCE from HBM and re-read from L2).
// Operands are the split-fp16 NHWC activations of the conv stack (float = hi + lo), read through one TMA tensor map
// per parity class (pixel strides doubled), so no thread ever computes an operand address; products are
// A_hi*B_hi + A_lo*B_hi + A_hi*B_lo (fp32-grade, FT_PREC_SPLIT) or A_hi*B_hi (FT_PREC_FAST), accumulated in TMEM.
//
// CTA = 6 warps, persistent over work units (a tile, or a part of a tile's chunk list):
//   warp 5  TMA   : the unit's A operand (C/64 k-blocks x hi/lo = up to 128 KiB) is loaded once and stays resident;
//                   the B chunks stream through a ring of 16 KiB stages (one plane of one k-block each, 5 stages in
//                   split mode, 9 in fast mode), so the L2 -> shared-memory traffic is about half of "A and B per chunk"
//   warp 4  MMA   : one thread, tcgen05.mma.cta_group::1.kind::f16 M=128 N=128 K=16; one TMEM accumulator stage
//                   (128 columns) per chunk, 4 stages in flight; frees A k-block kb right after its last use so the
//                   next unit's A loads overlap the tail of this one
//   warps 0-3 epilogue: warp q owns TMEM lanes 32q.. (= in1' row a0+q, lane = column b).  Per chunk row:
//                   tcgen05.ld 32 columns (one in2' row) -> per-warp smem slab S[b][b2] -> lane b reads the band
//                   S[b][b+ti], scales by 1/C, LeakyReLU, and stores out[n, ch, 2a+py, 2b+px] (fp32 NCHW).
#include <cuda.h>
#include <stdlib.h>

#include "ft_tc.cuh"

namespace ft {

// "This thread has read its part of a TMEM accumulator stage": a pure signal -- the reads are complete when tcgen05.wait::ld
// returns and no generic-proxy data is handed over -- so the arrive is RELAXED: it does not have to order the global stores the
// thread still has in flight (the previous rows of the band) before it.  (Same SASS and same time as the default .release.cta
// arrive on the boxes measured; it was tried while looking for the cluster-launch store slowdown described at the variant choice
// below, which it is not the cause of.)
__device__ __forceinline__ void tmem_stage_release(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}


constexpr int CD = 10;                 // displacement radius in parity-class pixels (max_displacement 20 / stride2 2)
constexpr int CDS = 2 * CD + 1;        // 21
constexpr int CORR_BM = 128, CORR_BN = 128, CORR_BK = 64;
constexpr int CORR_TILE = CORR_BM * CORR_BK * 2;      // 16 KiB: one operand plane of one k-block
constexpr int CORR_ROWS = 4, CORR_COLS = 32;          // pixel box of a tile / chunk
constexpr int CORR_KB_MAX = 4;                        // C <= 256
constexpr int CORR_EPI_WARPS = 4;
constexpr int CORR_THREADS = (CORR_EPI_WARPS + 2) * 32;
constexpr int CORR_ACC_STAGES = 4;                    // x 128 TMEM columns = all 512
constexpr int CORR_SLAB = 32 * 34 * 4;                // per epilogue warp: S[c][b] at c*34 + b: conflict-free writes (fixed c) and band reads (c = b + ti)

template <bool SPLIT>
struct CorrGeo {
    static constexpr int PLANES = SPLIT ? 2 : 1;
    static constexpr int A_BYTES = CORR_KB_MAX * PLANES * CORR_TILE;     // the A operand of a work unit stays resident
    static constexpr int SLABS = CORR_EPI_WARPS * CORR_SLAB;
    static constexpr int STAGES = (232448 - 1024 - 512 - A_BYTES - SLABS) / CORR_TILE;   // B ring: one plane of one k-block per stage
    static constexpr int SMEM = 1024 + A_BYTES + STAGES * CORR_TILE + SLABS + 512;
    static_assert(STAGES >= 4 && SMEM <= 232448, "shared memory budget");
};

struct alignas(64) CorrK {
    CUtensorMap a_map[4][2];     // [parity class][hi, lo] of in1
    CUtensorMap b_map[4][2];     // of in2 (4-row chunk boxes)
    CUtensorMap b2_map[4][2];    // of in2, 2-row boxes: the half chunk a CTA of a pair stages
    float* out;
    int B, H, W, C, kb;
    int a_tiles, num_tiles, parts, num_units;
    int b_tiles, ncc;   // column tiles of 32 in1' columns; in2' column chunks per row chunk: 1 (W/2 <= 32, chunk = columns 0..31) or 2 (wide
                        // frames: columns b0-10 .. b0+53 of the tile at b0, so that every lane's 21-wide band lies inside the two chunks)
    float inv;
    int leaky;
    int planar;     // output layout: 0 = fp32 NCHW [B,441,H,W] (the reference's), 1 = parity-planar [B,441,4,ceil(H/2),ceil(W/2)] (see ft_correlation_act)
    int dbg_skip;   // profiling knob (FT_DEV_SKIP): 1 = no A loads, 2 = no B loads, 4 = no MMAs, 8 = no global stores, 16 = no TMEM reads / staging (results are garbage)
};

// A work unit = part `part` of the chunk list of tile (n, cls, a0): chunks [c_lo, c_lo + nch) of in2' rows r_lo + 4*chunk.
struct CorrUnit {
    int n, cls, py, px, a0, H2c, r_lo, c_lo, nch, part;
};
__device__ __forceinline__ CorrUnit corr_unit(const CorrK& p, int unit) {
    CorrUnit t;
    const int tile = unit / p.parts;
    t.part = unit - tile * p.parts;
    const int at = tile % p.a_tiles, r = tile / p.a_tiles;     // consecutive CTAs share the in2' sub-image through L2
    t.cls = r & 3, t.n = r >> 2;
    t.py = t.cls >> 1, t.px = t.cls & 1;
    t.a0 = at * CORR_ROWS;
    t.H2c = (p.H - t.py + 1) >> 1;
    t.r_lo = max(0, t.a0 - CD);
    const int r_end = min(t.H2c, t.a0 + CORR_ROWS + CD);
    const int chunks = r_end > t.r_lo ? (r_end - t.r_lo + CORR_ROWS - 1) / CORR_ROWS : 0;
    const int per = (chunks + p.parts - 1) / p.parts;
    t.c_lo = min(chunks, t.part * per);
    t.nch = min(chunks, t.c_lo + per) - t.c_lo;
    return t;
}

template <bool SPLIT>
__global__ void __launch_bounds__(CORR_THREADS, 1) corr_tc_kernel(const __grid_constant__ CorrK p) {
    using G = CorrGeo<SPLIT>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t a_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t ring = a_base + G::A_BYTES;
    const uint32_t slabs = ring + G::STAGES * CORR_TILE;
    const uint32_t bar_base = slabs + G::SLABS;
    auto bfull_bar = [&](int s) { return bar_base + 8u * s; };
    auto bempty_bar = [&](int s) { return bar_base + 8u * (G::STAGES + s); };
    auto afull_bar = [&](int k) { return bar_base + 8u * (2 * G::STAGES + k); };
    auto aempty_bar = [&](int k) { return bar_base + 8u * (2 * G::STAGES + CORR_KB_MAX + k); };
    auto accf_bar = [&](int a) { return bar_base + 8u * (2 * G::STAGES + 2 * CORR_KB_MAX + a); };
    auto acce_bar = [&](int a) { return bar_base + 8u * (2 * G::STAGES + 2 * CORR_KB_MAX + CORR_ACC_STAGES + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * G::STAGES + 2 * CORR_KB_MAX + 2 * CORR_ACC_STAGES);
    static_assert(8 * (2 * G::STAGES + 2 * CORR_KB_MAX + 2 * CORR_ACC_STAGES + 1) <= 512, "barrier block");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < G::STAGES; ++s) {
            mbar_init(bfull_bar(s), 1);
            mbar_init(bempty_bar(s), 1);
        }
        for (int k = 0; k < CORR_KB_MAX; ++k) {
            mbar_init(afull_bar(k), 1);
            mbar_init(aempty_bar(k), 1);
        }
        for (int a = 0; a < CORR_ACC_STAGES; ++a) {
            mbar_init(accf_bar(a), 1);
            mbar_init(acce_bar(a), CORR_EPI_WARPS * 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == CORR_EPI_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == CORR_EPI_WARPS + 1) {
        // =========================== TMA producer (whole warp, one elected lane issues) ===========================
        {
            const bool ldA = !(p.dbg_skip & 1), ldB = !(p.dbg_skip & 2);
            int bit = 0, ut = 0;     // B ring stages issued / work units with chunks (phase of the A barriers)
            for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
                const CorrUnit t = corr_unit(p, unit);
                if (t.nch == 0) continue;
                for (int j = 0; j < t.nch; ++j) {
                    const int r0 = t.r_lo + (t.c_lo + j) * CORR_ROWS;
                    for (int kb = 0; kb < p.kb; ++kb) {
                        if (j == 0) {
                            // A k-block kb of this unit: its slot is freed by the MMA warp right after the previous unit's last use,
                            // and the B stages of the unit's first chunk follow it immediately, so the unit switch costs no bubble
                            mbar_wait(aempty_bar(kb), (ut & 1) ^ 1);
                            const uint32_t dst = a_base + kb * G::PLANES * CORR_TILE;
                            if (elect_one()) {
                                mbar_arrive_expect_tx(afull_bar(kb), ldA ? (uint32_t)(G::PLANES * CORR_TILE) : 0u);
                                if (ldA) {
                                    tma_load_4d(dst, &p.a_map[t.cls][0], afull_bar(kb), kb * CORR_BK, 0, t.a0, t.n);
                                    if (SPLIT) tma_load_4d(dst + CORR_TILE, &p.a_map[t.cls][1], afull_bar(kb), kb * CORR_BK, 0, t.a0, t.n);
                                }
                            }
                            __syncwarp();
                        }
#pragma unroll
                        for (int pl = 0; pl < G::PLANES; ++pl, ++bit) {
                            const int s = bit % G::STAGES;
                            mbar_wait(bempty_bar(s), ((bit / G::STAGES) & 1) ^ 1);
                            if (elect_one()) {
                                mbar_arrive_expect_tx(bfull_bar(s), ldB ? (uint32_t)CORR_TILE : 0u);
                                if (ldB) tma_load_4d(ring + s * CORR_TILE, &p.b_map[t.cls][pl], bfull_bar(s), kb * CORR_BK, 0, r0, t.n);
                            }
                            __syncwarp();
                        }
                    }
                }
                ++ut;
            }
        }
    } else if (warp == CORR_EPI_WARPS) {
        // =========================== MMA issuer ===========================
        // The whole warp runs the loop with warp-uniform values; one elected lane issues the tcgen05 instructions (in single-lane
        // code the compiler wraps every UTCHMMA in an ELECT + R2UR.BROADCAST sequence of ~100 cycles, slower than the MMA itself).
        {
            constexpr uint32_t idesc = f16_idesc(CORR_BM, CORR_BN);
            const bool do_mma = !(p.dbg_skip & 4);
            int bit = 0, ut = 0, acc_it = 0;
            for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
                const CorrUnit t = corr_unit(p, unit);
                if (t.nch == 0) continue;
                for (int j = 0; j < t.nch; ++j, ++acc_it) {
                    const int a = acc_it % CORR_ACC_STAGES;
                    mbar_wait(acce_bar(a), ((acc_it / CORR_ACC_STAGES) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(a * CORR_BN);
                    for (int kb = 0; kb < p.kb; ++kb) {
                        if (j == 0) mbar_wait(afull_bar(kb), ut & 1);
                        const uint32_t a_hi = a_base + kb * G::PLANES * CORR_TILE;
                        {   // B_hi: A_hi*B_hi (+ A_lo*B_hi)
                            const int s = bit % G::STAGES;
                            mbar_wait(bfull_bar(s), (bit / G::STAGES) & 1);
                            tc_fence_after();
                            const uint32_t b = ring + s * CORR_TILE;
                            if (elect_one()) {
#pragma unroll
                                for (int k = 0; k < CORR_BK / 16; ++k) {
                                    if (!do_mma) break;
                                    const uint64_t db = sw128_desc(b + k * 32);
                                    tc_mma_f16(d_tmem, sw128_desc(a_hi + k * 32), db, idesc, (kb > 0) || (k != 0));
                                    if (SPLIT) tc_mma_f16(d_tmem, sw128_desc(a_hi + CORR_TILE + k * 32), db, idesc, 1);
                                }
                                tc_commit(bempty_bar(s));
                                if (!SPLIT) {
                                    if (j == t.nch - 1) tc_commit(aempty_bar(kb));
                                    if (kb == p.kb - 1) tc_commit(accf_bar(a));
                                }
                            }
                            __syncwarp();
                            ++bit;
                        }
                        if (SPLIT) {   // B_lo: A_hi*B_lo
                            const int s = bit % G::STAGES;
                            mbar_wait(bfull_bar(s), (bit / G::STAGES) & 1);
                            tc_fence_after();
                            const uint32_t b = ring + s * CORR_TILE;
                            if (elect_one()) {
#pragma unroll
                                for (int k = 0; k < CORR_BK / 16; ++k) {
                                    if (!do_mma) break;
                                    tc_mma_f16(d_tmem, sw128_desc(a_hi + k * 32), sw128_desc(b + k * 32), idesc, 1);
                                }
                                tc_commit(bempty_bar(s));
                                if (j == t.nch - 1) tc_commit(aempty_bar(kb));     // the unit's last use of this A k-block
                                if (kb == p.kb - 1) tc_commit(accf_bar(a));
                            }
                            __syncwarp();
                            ++bit;
                        }
                    }
                }
                ++ut;
            }
        }
    } else {
        // =========================== epilogue (warps 0-3) ===========================
        const int q = warp;                                       // TMEM lane quarter = in1' row a0 + q; lane = column b
        float* S = reinterpret_cast<float*>(smem_raw + (slabs - smem_u32(smem_raw))) + warp * (32 * 34);
        // NCHW: lane b writes column 2b+px of row 2a+py; parity-planar: class plane [H2, W2], lane b writes column b (see the pair kernel)
        const int H2 = (p.H + 1) >> 1, W2 = (p.W + 1) >> 1;
        const size_t plane = p.planar ? (size_t)4 * H2 * W2 : (size_t)p.H * p.W;
        int acc_it = 0;
        for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
            const CorrUnit t = corr_unit(p, unit);
            const int a = t.a0 + q;
            const int y = 2 * a + t.py, x = 2 * lane + t.px;
            const bool valid = y < p.H && x < p.W && !(p.dbg_skip & 8);
            float* o = p.out + (size_t)t.n * (CDS * CDS) * plane + (p.planar ? ((size_t)t.cls * H2 + a) * W2 + lane : (size_t)y * p.W + x);
            if (t.part == 0) {
                // displacement rows that fall outside the image: the correlation is zero there (never visited by a chunk)
                for (int tj = -CD; tj <= CD; ++tj) {
                    const int row = a + tj;
                    if (row >= 0 && row < t.H2c) continue;
                    if (valid) {
#pragma unroll
                        for (int ti = 0; ti < CDS; ++ti) o[(size_t)((tj + CD) * CDS + ti) * plane] = 0.f;
                    }
                }
            }
            for (int j = 0; j < t.nch; ++j, ++acc_it) {
                const int as = acc_it % CORR_ACC_STAGES;
                const int r0 = t.r_lo + (t.c_lo + j) * CORR_ROWS;
                mbar_wait(accf_bar(as), (acc_it / CORR_ACC_STAGES) & 1);
                tc_fence_after();
                const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * CORR_BN);
#pragma unroll
                for (int h = 0; h < 2; ++h) {                      // chunk rows 2h, 2h+1
                    uint32_t v0[32], v1[32];
                    const int row0 = r0 + 2 * h, row1 = row0 + 1;
                    const int tj0 = row0 - a, tj1 = row1 - a;
                    const bool use0 = tj0 >= -CD && tj0 <= CD && row0 < t.H2c && !(p.dbg_skip & 16);     // warp-uniform
                    const bool use1 = tj1 >= -CD && tj1 <= CD && row1 < t.H2c && !(p.dbg_skip & 16);
                    if (use0) TMEM_LD_32(t_row + (2 * h) * 32, v0);
                    if (use1) TMEM_LD_32(t_row + (2 * h + 1) * 32, v1);
                    if (use0 || use1) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (h == 1) {
                        tc_fence_before();
                        tmem_stage_release(map_to_cta(acce_bar(as), 0));   // this thread is done with the TMEM stage (own CTA = cluster rank 0)
                    }
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr) {
                        const bool use = rr ? use1 : use0;
                        if (!use) continue;
                        const int tj = rr ? tj1 : tj0;
                        __syncwarp();
#pragma unroll
                        for (int c = 0; c < 32; ++c) S[c * 34 + lane] = __uint_as_float(rr ? v1[c] : v0[c]);
                        __syncwarp();
                        if (valid) {
                            float* oc = o + (size_t)((tj + CD) * CDS) * plane;
#pragma unroll
                            for (int ti = 0; ti < CDS; ++ti) {
                                const int b2 = lane + ti - CD;
                                float val = (b2 >= 0 && b2 < 32) ? S[b2 * 34 + lane] : 0.f;   // columns >= W/2 were zero-filled by TMA
                                val *= p.inv;
                                if (p.leaky) val = val > 0.f ? val : 0.1f * val;
                                oc[(size_t)ti * plane] = val;
                            }
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == CORR_EPI_WARPS) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// Re-layout for the reference-signature entry point: contiguous fp32 NCHW -> split-fp16 NHWC planes (hi, lo).  A block
// transposes a tile of 64 channels x 32 pixels through shared memory: 128-byte coalesced reads along x, 16-byte stores
// that complete full 128-byte lines of a pixel's channel vector.
__global__ void __launch_bounds__(256) corr_relayout_kernel(const float* __restrict__ src, __half* __restrict__ hi, __half* __restrict__ lo, int C, long long npix) {
    __shared__ float tile[64][33];
    const long long p0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 64, n = blockIdx.z;
    const float* s = src + (size_t)n * C * npix;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = c0 + w * 8 + i;
        const long long px = p0 + lane;
        tile[w * 8 + i][lane] = (c < C && px < npix) ? __ldg(s + (size_t)c * npix + px) : 0.f;
    }
    __syncthreads();
    const int pl = threadIdx.x >> 3, g = threadIdx.x & 7;       // pixel of the tile, 8-channel group
    const long long px = p0 + pl;
    if (px < npix && c0 + g * 8 < C) {
        Half8 h8, l8;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            __half h0, l0, h1, l1;
            split_f32(tile[g * 8 + 2 * k][pl], h0, l0);
            split_f32(tile[g * 8 + 2 * k + 1][pl], h1, l1);
            h8.v[k] = __halves2half2(h0, h1), l8.v[k] = __halves2half2(l0, l1);
        }
        const size_t off = ((size_t)n * npix + px) * C + c0 + g * 8;
        *reinterpret_cast<Half8*>(hi + off) = h8;
        *reinterpret_cast<Half8*>(lo + off) = l8;
    }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): the two CTAs of a cluster (the two SMs of a TPC) take two adjacent in1' tiles
// (8 rows = 256 pixels, M = 256) of the same parity class; each keeps ITS tile resident and stages only HALF of every
// in2' chunk (2 of its 4 rows), the tensor cores of both SMs read both halves.  Per SM a chunk then costs 64 KiB of
// shared-memory fill instead of 128 KiB against the same 3072 tensor-pipe cycles -- the single-CTA main loop is bound by
// exactly that fill (ring depth x stage latency).  The union of the two tiles' displacement windows is 28 in2' rows:
// 16 pair-chunks per 24-row class instead of 2 x 15 tile-chunks (+7 % MMA work).
// Protocol as in conv_tc_kernel<PAIR>: rank 0 issues every MMA; the peer's MMA warp relays "my stage is full"; commits are
// multicast to both CTAs; the leader's acc_empty barriers count the epilogue threads of both CTAs.
// ------------------------------------------------------------------------------------------------
constexpr int CP_HALF = CORR_TILE / 2;                // 8 KiB: 64 in2' pixels x 64 channels, one plane

template <bool SPLIT>
struct CorrPairGeo {
    static constexpr int PLANES = SPLIT ? 2 : 1;
    static constexpr int A_BYTES = CORR_KB_MAX * PLANES * CORR_TILE;
    static constexpr int SLABS = CORR_EPI_WARPS * CORR_SLAB;
    static constexpr int STAGE = PLANES * CP_HALF;                    // one k-block of this CTA's half chunk, hi (and lo)
    static constexpr int STAGES_FIT = (232448 - 1024 - 512 - A_BYTES - SLABS) / STAGE;
    static constexpr int STAGES = STAGES_FIT > 10 ? 10 : STAGES_FIT;
    static constexpr int SMEM = 1024 + A_BYTES + STAGES * STAGE + SLABS + 512;
    static_assert(STAGES >= 4 && SMEM <= 232448, "shared memory budget");
};

struct CorrPairUnit {
    int n, cls, py, px, a0, H2c, r_lo, c_lo, nch, part, b0;
};
__device__ __forceinline__ CorrPairUnit corr_pair_unit(const CorrK& p, int unit, int rank) {
    CorrPairUnit t;
    const int ap_tiles = (p.a_tiles + 1) >> 1;
    const int tile = unit / p.parts;
    t.part = unit - tile * p.parts;
    const int ap = tile % ap_tiles;
    int r = tile / ap_tiles;
    t.b0 = (r % p.b_tiles) * CORR_COLS;
    r /= p.b_tiles;
    t.cls = r & 3, t.n = r >> 2;
    t.py = t.cls >> 1, t.px = t.cls & 1;
    const int A0 = ap * 2 * CORR_ROWS;
    t.a0 = A0 + rank * CORR_ROWS;                                  // this CTA's in1' rows (may lie beyond the image: all padding)
    t.H2c = (p.H - t.py + 1) >> 1;
    t.r_lo = max(0, A0 - CD);
    const int r_end = min(t.H2c, A0 + 2 * CORR_ROWS + CD);
    const int chunks = (r_end > t.r_lo ? (r_end - t.r_lo + CORR_ROWS - 1) / CORR_ROWS : 0) * p.ncc;    // (row chunk, column chunk) pairs, column fastest
    const int per = (chunks + p.parts - 1) / p.parts;
    t.c_lo = min(chunks, t.part * per);
    t.nch = min(chunks, t.c_lo + per) - t.c_lo;
    return t;
}
// chunk c of a unit -> first in2' row of the row chunk and first in2' column of the column chunk
__device__ __forceinline__ void corr_pair_chunk(const CorrK& p, const CorrPairUnit& t, int c, int& r0, int& c0) {
    const int j = p.ncc == 2 ? c >> 1 : c, x = p.ncc == 2 ? c & 1 : 0;
    r0 = t.r_lo + j * CORR_ROWS;
    c0 = p.ncc == 2 ? t.b0 - CD + CORR_COLS * x : 0;
}

template <bool SPLIT>
__global__ void __launch_bounds__(CORR_THREADS, 1) corr_pair_kernel(const __grid_constant__ CorrK p) {
    using G = CorrPairGeo<SPLIT>;
    const int rank = (int)cluster_ctarank();
    const int first_unit = (int)(blockIdx.x >> 1), unit_step = (int)(gridDim.x >> 1);
    extern __shared__ uint8_t smem_raw[];
    const uint32_t a_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t ring = a_base + G::A_BYTES;
    const uint32_t slabs = ring + G::STAGES * G::STAGE;
    const uint32_t bar_base = slabs + G::SLABS;
    // barrier slots: bfull[S] | bempty[S] | afull[4] | aempty[4] | accf[4] | acce[4] | tmem ptr | peer_bfull[S] | peer_afull[4]
    auto bfull_bar = [&](int s) { return bar_base + 8u * s; };
    auto bempty_bar = [&](int s) { return bar_base + 8u * (G::STAGES + s); };
    auto afull_bar = [&](int k) { return bar_base + 8u * (2 * G::STAGES + k); };
    auto aempty_bar = [&](int k) { return bar_base + 8u * (2 * G::STAGES + CORR_KB_MAX + k); };
    auto accf_bar = [&](int a) { return bar_base + 8u * (2 * G::STAGES + 2 * CORR_KB_MAX + a); };
    auto acce_bar = [&](int a) { return bar_base + 8u * (2 * G::STAGES + 2 * CORR_KB_MAX + CORR_ACC_STAGES + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * G::STAGES + 2 * CORR_KB_MAX + 2 * CORR_ACC_STAGES);
    auto peer_bfull_bar = [&](int s) { return bar_base + 8u * (2 * G::STAGES + 2 * CORR_KB_MAX + 2 * CORR_ACC_STAGES + 1 + s); };
    auto peer_afull_bar = [&](int k) { return bar_base + 8u * (3 * G::STAGES + 2 * CORR_KB_MAX + 2 * CORR_ACC_STAGES + 1 + k); };
    static_assert(8 * (3 * G::STAGES + 3 * CORR_KB_MAX + 2 * CORR_ACC_STAGES + 1) <= 512, "barrier block");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < G::STAGES; ++s) {
            mbar_init(bfull_bar(s), 1);
            mbar_init(bempty_bar(s), 1);
            mbar_init(peer_bfull_bar(s), 1);
        }
        for (int k = 0; k < CORR_KB_MAX; ++k) {
            mbar_init(afull_bar(k), 1);
            mbar_init(aempty_bar(k), 1);
            mbar_init(peer_afull_bar(k), 1);
        }
        for (int a = 0; a < CORR_ACC_STAGES; ++a) {
            mbar_init(accf_bar(a), 1);
            mbar_init(acce_bar(a), 2 * CORR_EPI_WARPS * 32);          // the epilogue threads of both CTAs release the stage
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == CORR_EPI_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == CORR_EPI_WARPS + 1) {
        // =========================== TMA producer: this CTA's A tile and its half (2 rows) of every chunk ===========================
        const bool ldA = !(p.dbg_skip & 1), ldB = !(p.dbg_skip & 2);
        int bit = 0, ut = 0;
        for (int unit = first_unit; unit < p.num_units; unit += unit_step) {
            const CorrPairUnit t = corr_pair_unit(p, unit, rank);
            if (t.nch == 0) continue;
            for (int j = 0; j < t.nch; ++j) {
                int r0, c0;
                corr_pair_chunk(p, t, t.c_lo + j, r0, c0);
                r0 += 2 * rank;
                for (int kb = 0; kb < p.kb; ++kb, ++bit) {
                    if (j == 0) {
                        mbar_wait_cluster(aempty_bar(kb), (ut & 1) ^ 1);
                        const uint32_t dst = a_base + kb * G::PLANES * CORR_TILE;
                        if (elect_one()) {
                            mbar_arrive_expect_tx(afull_bar(kb), ldA ? (uint32_t)(G::PLANES * CORR_TILE) : 0u);
                            if (ldA) {
                                tma_load_4d(dst, &p.a_map[t.cls][0], afull_bar(kb), kb * CORR_BK, t.b0, t.a0, t.n);
                                if (SPLIT) tma_load_4d(dst + CORR_TILE, &p.a_map[t.cls][1], afull_bar(kb), kb * CORR_BK, t.b0, t.a0, t.n);
                            }
                        }
                        __syncwarp();
                    }
                    const int s = bit % G::STAGES;
                    mbar_wait_cluster(bempty_bar(s), ((bit / G::STAGES) & 1) ^ 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(bfull_bar(s), ldB ? (uint32_t)G::STAGE : 0u);
                        if (ldB) {       // columns left of 0 / right of the sub-image are zero-filled by the TMA unit
                            tma_load_4d(ring + s * G::STAGE, &p.b2_map[t.cls][0], bfull_bar(s), kb * CORR_BK, c0, r0, t.n);
                            if (SPLIT) tma_load_4d(ring + s * G::STAGE + CP_HALF, &p.b2_map[t.cls][1], bfull_bar(s), kb * CORR_BK, c0, r0, t.n);
                        }
                    }
                    __syncwarp();
                }
            }
            ++ut;
        }
    } else if (warp == CORR_EPI_WARPS) {
        if (rank == 0) {
            // =========================== MMA issuer (leader CTA; whole warp, one elected lane issues) ===========================
            constexpr uint32_t idesc = f16_idesc(2 * CORR_BM, CORR_BN);
            const bool do_mma = !(p.dbg_skip & 4);
            int bit = 0, ut = 0, acc_it = 0;
            for (int unit = first_unit; unit < p.num_units; unit += unit_step) {
                const CorrPairUnit t = corr_pair_unit(p, unit, 0);
                if (t.nch == 0) continue;
                for (int j = 0; j < t.nch; ++j, ++acc_it) {
                    const int a = acc_it % CORR_ACC_STAGES;
                    mbar_wait_cluster(acce_bar(a), ((acc_it / CORR_ACC_STAGES) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(a * CORR_BN);
                    for (int kb = 0; kb < p.kb; ++kb, ++bit) {
                        if (j == 0) {
                            mbar_wait(afull_bar(kb), ut & 1);
                            mbar_wait_cluster(peer_afull_bar(kb), ut & 1);
                        }
                        const uint32_t a_hi = a_base + kb * G::PLANES * CORR_TILE;
                        const int s = bit % G::STAGES;
                        mbar_wait(bfull_bar(s), (bit / G::STAGES) & 1);
                        mbar_wait_cluster(peer_bfull_bar(s), (bit / G::STAGES) & 1);
                        tc_fence_after();
                        const uint32_t b = ring + s * G::STAGE;
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < CORR_BK / 16; ++k) {
                                if (!do_mma) break;
                                const uint64_t da = sw128_desc(a_hi + k * 32), db = sw128_desc(b + k * 32);
                                tc_mma_f16_pair(d_tmem, da, db, idesc, (kb > 0) || (k != 0));                                   // A_hi * B_hi
                                if (SPLIT) {
                                    tc_mma_f16_pair(d_tmem, sw128_desc(a_hi + CORR_TILE + k * 32), db, idesc, 1);               // A_lo * B_hi
                                    tc_mma_f16_pair(d_tmem, da, sw128_desc(b + CP_HALF + k * 32), idesc, 1);                    // A_hi * B_lo
                                }
                            }
                            tc_commit_pair(bempty_bar(s));
                            if (j == t.nch - 1) tc_commit_pair(aempty_bar(kb));
                            if (kb == p.kb - 1) tc_commit_pair(accf_bar(a));
                        }
                        __syncwarp();
                    }
                }
                ++ut;
            }
        } else if (lane == 0) {
            // peer CTA: tell the leader when this CTA's A k-blocks / B half stages have landed (same order as the leader consumes them)
            int bit = 0, ut = 0;
            for (int unit = first_unit; unit < p.num_units; unit += unit_step) {
                const CorrPairUnit t = corr_pair_unit(p, unit, 1);
                if (t.nch == 0) continue;
                for (int j = 0; j < t.nch; ++j)
                    for (int kb = 0; kb < p.kb; ++kb, ++bit) {
                        if (j == 0) {
                            mbar_wait(afull_bar(kb), ut & 1);
                            mbar_arrive_remote(map_to_cta(peer_afull_bar(kb), 0));
                        }
                        const int s = bit % G::STAGES;
                        mbar_wait(bfull_bar(s), (bit / G::STAGES) & 1);
                        mbar_arrive_remote(map_to_cta(peer_bfull_bar(s), 0));
                    }
                ++ut;
            }
        }
    } else {
        // =========================== epilogue (warps 0-3): this CTA's 128 TMEM lanes = its in1' tile ===========================
        const int q = warp;
        float* S = reinterpret_cast<float*>(smem_raw + (slabs - smem_u32(smem_raw))) + warp * (32 * 34);
        // NCHW: lane b writes column 2b+px of row 2a+py (4-byte stores 8 bytes apart: half-filled sectors).  Parity-planar: every
        // class owns a dense [H2, W2] plane per channel, lane b writes column b: 128 contiguous bytes per warp store.
        const int H2 = (p.H + 1) >> 1, W2 = (p.W + 1) >> 1;
        const size_t plane = p.planar ? (size_t)4 * H2 * W2 : (size_t)p.H * p.W;
        const uint32_t acce_remote0 = map_to_cta(acce_bar(0), 0);
        int acc_it = 0;
        for (int unit = first_unit; unit < p.num_units; unit += unit_step) {
            const CorrPairUnit t = corr_pair_unit(p, unit, rank);
            const int a = t.a0 + q;
            const int bcol = t.b0 + lane;                                          // this lane's in1' column
            const int y = 2 * a + t.py, x = 2 * bcol + t.px;
            const bool valid = y < p.H && x < p.W && !(p.dbg_skip & 8);
            float* o = p.out + (size_t)t.n * (CDS * CDS) * plane + (p.planar ? ((size_t)t.cls * H2 + a) * W2 + bcol : (size_t)y * p.W + x);
            if (t.part == 0) {
                for (int tj = -CD; tj <= CD; ++tj) {              // displacement rows outside the image: zeros
                    const int row = a + tj;
                    if (row >= 0 && row < t.H2c) continue;
                    if (valid) {
#pragma unroll
                        for (int ti = 0; ti < CDS; ++ti) o[(size_t)((tj + CD) * CDS + ti) * plane] = 0.f;
                    }
                }
            }
            for (int j = 0; j < t.nch; ++j, ++acc_it) {
                const int as = acc_it % CORR_ACC_STAGES;
                int r0, c0;
                corr_pair_chunk(p, t, t.c_lo + j, r0, c0);
                const int cshift = p.ncc == 2 ? c0 - t.b0 + CD : 0;               // wide: 0 or 32 (the chunk's first band position)
                mbar_wait_cluster(accf_bar(as), (acc_it / CORR_ACC_STAGES) & 1);
                tc_fence_after();
                const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * CORR_BN);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t v0[32], v1[32];
                    const int row0 = r0 + 2 * h, row1 = row0 + 1;
                    const int tj0 = row0 - a, tj1 = row1 - a;
                    const bool use0 = tj0 >= -CD && tj0 <= CD && row0 < t.H2c && !(p.dbg_skip & 16);     // warp-uniform
                    const bool use1 = tj1 >= -CD && tj1 <= CD && row1 < t.H2c && !(p.dbg_skip & 16);
                    if (use0) TMEM_LD_32(t_row + (2 * h) * 32, v0);
                    if (use1) TMEM_LD_32(t_row + (2 * h + 1) * 32, v1);
                    if (use0 || use1) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (h == 1) {
                        tc_fence_before();
                        tmem_stage_release(acce_remote0 + 8u * as);      // the leader's barrier counts both CTAs' epilogue threads
                    }
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr) {
                        const bool use = rr ? use1 : use0;
                        if (!use) continue;
                        const int tj = rr ? tj1 : tj0;
                        __syncwarp();
#pragma unroll
                        for (int c = 0; c < 32; ++c) S[c * 34 + lane] = __uint_as_float(rr ? v1[c] : v0[c]);
                        __syncwarp();
                        if (valid) {
                            float* oc = o + (size_t)((tj + CD) * CDS) * plane;
#pragma unroll
                            for (int ti = 0; ti < CDS; ++ti) {
                                // narrow: chunk column = in2' column, b2 = b + ti - 10 (outside the chunk = outside the image: zero).
                                // wide: the chunk starts at in2' column b0 - 10 + cshift, so b2 = lane + ti - cshift; a band position
                                // outside this chunk belongs to the other column chunk of the row and is not written here.
                                const int b2 = p.ncc == 2 ? lane + ti - cshift : lane + ti - CD;
                                const bool inside = b2 >= 0 && b2 < 32;
                                if (p.ncc == 2 && !inside) continue;
                                float val = inside ? S[b2 * 34 + lane] : 0.f;
                                val *= p.inv;
                                if (p.leaky) val = val > 0.f ? val : 0.1f * val;
                                oc[(size_t)ti * plane] = val;
                            }
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // neither CTA may leave (or free TMEM) while its peer can still touch its smem / barriers
    if (warp == CORR_EPI_WARPS) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

template <bool SPLIT>
static int launch_corr_pair(const CorrK& k0, cudaStream_t s) {
    using G = CorrPairGeo<SPLIT>;
    static bool attr_set[MAX_DEVICES];
    if (ensure_dyn_smem(corr_pair_kernel<SPLIT>, G::SMEM, attr_set, "corr_pair_kernel")) return 1;
    CorrK k = k0;
    k.num_units = k.B * 4 * ((k.a_tiles + 1) / 2) * k.b_tiles * k.parts;
    const int max_clusters = device_sms() / 2;
    const int clusters = k.num_units < max_clusters ? k.num_units : max_clusters;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters), cfg.blockDim = dim3(CORR_THREADS), cfg.dynamicSmemBytes = G::SMEM, cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, corr_pair_kernel<SPLIT>, k);
    if (e != cudaSuccess) {
        set_error("corr_pair_kernel<%d>: cluster launch failed: %s", (int)SPLIT, cudaGetErrorString(e));
        return 1;
    }
    return check_launch("corr_pair_kernel");
}

// ------------------------------------------------------------------------------------------------
// Variant 3: the A operand lives in TENSOR MEMORY.  A unit's in1' tile (C/64 k-blocks x hi/lo) is staged through the same
// shared-memory ring as the in2' chunks and copied into TMEM columns [256, 512) by tcgen05.cp (issued by the MMA thread;
// tcgen05.cp and tcgen05.mma execute in issue order, so no barrier is needed between the copy and the MMAs that read it,
// nor between a unit's last MMA and the next unit's copy).  Shared memory then holds nothing but the ring: 5 stages of
// 32 KiB (one k-block, hi and lo plane) = 160 KiB of operand bytes in flight instead of 80 KiB -- the main loop of variant 2 is bound by the latency of a
// ring stage times the number of stages that fit beside the resident A tile.  Accumulators: 2 TMEM stages x 128 columns
// (columns [0, 256)); 8 epilogue warps (two per lane quarter, two chunk rows each) release a stage right after their
// tcgen05.ld.
// ------------------------------------------------------------------------------------------------
// dev instrumentation (FT_DEV_SKIP bit 64): cycles each role spends waiting, per CTA: [role][slot]
__device__ long long g_corr_prof[192][16];   // dev instrumentation, one row per persistent CTA
#define MBAR_WAIT1(bar, par) do { if (p.dbg_skip & 32) mbar_wait(bar, par); else mbar_wait_spin(bar, par); } while (0)
#define TIMED(slot, stmt) do { if (prof) { const long long t_ = clock64(); stmt; pw[slot] += clock64() - t_; } else { stmt; } } while (0)
constexpr int C3_EPI_WARPS = 8;
constexpr int C3_THREADS = (C3_EPI_WARPS + 2) * 32;
constexpr int C3_ACC_STAGES = 2;
constexpr int C3_A_COL = 256;                         // first TMEM column of the A operand: [hi: kb*32 + k*8 | lo: 128 + ...]

template <bool SPLIT>
struct Corr3Geo {
    static constexpr int PLANES = SPLIT ? 2 : 1;
    static constexpr int SLABS = C3_EPI_WARPS * CORR_SLAB;
    static constexpr int STAGE = PLANES * CORR_TILE;                 // one 64-channel k-block of one operand, hi (and lo) plane
    static constexpr int STAGES = (232448 - 1024 - 512 - SLABS) / STAGE;
    static constexpr int SMEM = 1024 + STAGES * STAGE + SLABS + 512;
    static_assert(STAGES >= 4 && STAGES <= 16 && SMEM <= 232448, "shared memory budget");
};

__device__ __forceinline__ void tc_cp_128x256b(uint32_t d_tmem, uint64_t s_desc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(d_tmem), "l"(s_desc) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T
__device__ __forceinline__ void tc_mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

template <bool SPLIT>
__global__ void __launch_bounds__(C3_THREADS, 1) corr_tc3_kernel(const __grid_constant__ CorrK p) {
    using G = Corr3Geo<SPLIT>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t slabs = ring + G::STAGES * G::STAGE;
    const uint32_t bar_base = slabs + G::SLABS;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (G::STAGES + s); };
    auto accf_bar = [&](int a) { return bar_base + 8u * (2 * G::STAGES + a); };
    auto acce_bar = [&](int a) { return bar_base + 8u * (2 * G::STAGES + C3_ACC_STAGES + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * G::STAGES + 2 * C3_ACC_STAGES);
    static_assert(8 * (2 * G::STAGES + 2 * C3_ACC_STAGES + 1) <= 512, "barrier block");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool prof = (p.dbg_skip & 64) != 0;
    long long pw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long t_start = clock64();

    if (threadIdx.x == 0) {
        for (int s = 0; s < G::STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < C3_ACC_STAGES; ++a) {
            mbar_init(accf_bar(a), 1);
            mbar_init(acce_bar(a), C3_EPI_WARPS * 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == C3_EPI_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == C3_EPI_WARPS + 1) {
        // =========================== TMA producer: per unit the A stages, then the B stages of every chunk ===========================
        if (lane == 0) {
            const bool ldA = !(p.dbg_skip & 1), ldB = !(p.dbg_skip & 2);
            int it = 0;
            for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
                const CorrUnit t = corr_unit(p, unit);
                if (t.nch == 0) continue;
                for (int j = -1; j < t.nch; ++j) {                 // j = -1: the unit's A operand
                    const int r0 = j < 0 ? t.a0 : t.r_lo + (t.c_lo + j) * CORR_ROWS;
                    const CUtensorMap* maps = j < 0 ? p.a_map[t.cls] : p.b_map[t.cls];
                    const bool ld = j < 0 ? ldA : ldB;
                    for (int kb = 0; kb < p.kb; ++kb, ++it) {
                        const int s = it % G::STAGES;
                        TIMED(0, MBAR_WAIT1(empty_bar(s), ((it / G::STAGES) & 1) ^ 1));
                        mbar_arrive_expect_tx(full_bar(s), ld ? (uint32_t)G::STAGE : 0u);
#pragma unroll
                        for (int pl = 0; pl < G::PLANES; ++pl)
                            if (ld) tma_load_4d(ring + s * G::STAGE + pl * CORR_TILE, &maps[pl], full_bar(s), kb * CORR_BK, 0, r0, t.n);
                    }
                }
            }
        }
    } else if (warp == C3_EPI_WARPS) {
        // =========================== MMA issuer (also issues the smem -> TMEM copies of A) ===========================
        // The whole warp runs the loop with warp-uniform values; one elected lane issues the tcgen05 instructions.
        {
            constexpr uint32_t idesc = f16_idesc(CORR_BM, CORR_BN);
            const bool do_mma = !(p.dbg_skip & 4);
            const uint32_t a_tmem = tmem_base + C3_A_COL;
            int it = 0, acc_it = 0;
            for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
                const CorrUnit t = corr_unit(p, unit);
                if (t.nch == 0) continue;
                for (int kb = 0; kb < p.kb; ++kb, ++it) {
                    const int s = it % G::STAGES;
                    TIMED(0, MBAR_WAIT1(full_bar(s), (it / G::STAGES) & 1));
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int pl = 0; pl < G::PLANES; ++pl)
#pragma unroll
                            for (int k = 0; k < CORR_BK / 16; ++k)
                                tc_cp_128x256b(a_tmem + pl * 128 + kb * 32 + k * 8, sw128_desc(ring + s * G::STAGE + pl * CORR_TILE + k * 32));
                        tc_commit(empty_bar(s));
                    }
                    __syncwarp();
                }
                for (int j = 0; j < t.nch; ++j, ++acc_it) {
                    const int a = acc_it % C3_ACC_STAGES;
                    TIMED(1, MBAR_WAIT1(acce_bar(a), ((acc_it / C3_ACC_STAGES) & 1) ^ 1));
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(a * CORR_BN);
                    for (int kb = 0; kb < p.kb; ++kb, ++it) {
                        const int s = it % G::STAGES;
                        TIMED(2, MBAR_WAIT1(full_bar(s), (it / G::STAGES) & 1));
                        TIMED(4, tc_fence_after());
                        const uint32_t b = ring + s * G::STAGE;
                        const long long t_issue = prof ? clock64() : 0;
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < CORR_BK / 16; ++k) {
                                if (!do_mma) break;
                                const uint64_t db = sw128_desc(b + k * 32);
                                const uint32_t ah = a_tmem + kb * 32 + k * 8;
                                tc_mma_f16_ts(d_tmem, ah, db, idesc, (kb > 0) || (k != 0));                // A_hi * B_hi
                                if (SPLIT) {
                                    tc_mma_f16_ts(d_tmem, ah + 128, db, idesc, 1);                         // A_lo * B_hi
                                    tc_mma_f16_ts(d_tmem, ah, sw128_desc(b + CORR_TILE + k * 32), idesc, 1);   // A_hi * B_lo
                                }
                            }
                            const long long t_c = prof ? clock64() : 0;
                            tc_commit(empty_bar(s));
                            if (kb == p.kb - 1) tc_commit(accf_bar(a));
                            if (prof) pw[6] += clock64() - t_c;
                        }
                        __syncwarp();
                        if (prof) pw[5] += clock64() - t_issue;
                    }
                }
            }
        }
    } else {
        // =========================== epilogue (warps 0-7): warp (q, h) = TMEM lane quarter q, chunk rows 2h, 2h+1 ===========================
        const int q = warp & 3, h = warp >> 2;
        float* S = reinterpret_cast<float*>(smem_raw + (slabs - smem_u32(smem_raw))) + warp * (32 * 34);
        const size_t plane = (size_t)p.H * p.W;
        int acc_it = 0;
        for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
            const CorrUnit t = corr_unit(p, unit);
            const int a = t.a0 + q;
            const int y = 2 * a + t.py, x = 2 * lane + t.px;
            const bool valid = y < p.H && x < p.W && !(p.dbg_skip & 8);
            float* o = p.out + (size_t)t.n * (CDS * CDS) * plane + (size_t)y * p.W + x;
            if (t.part == 0) {
                for (int tj = -CD + h; tj <= CD; tj += 2) {       // displacement rows outside the image: zeros
                    const int row = a + tj;
                    if (row >= 0 && row < t.H2c) continue;
                    if (valid) {
#pragma unroll
                        for (int ti = 0; ti < CDS; ++ti) o[(size_t)((tj + CD) * CDS + ti) * plane] = 0.f;
                    }
                }
            }
            for (int j = 0; j < t.nch; ++j, ++acc_it) {
                const int as = acc_it % C3_ACC_STAGES;
                const int r0 = t.r_lo + (t.c_lo + j) * CORR_ROWS;
                TIMED(0, mbar_wait(accf_bar(as), (acc_it / C3_ACC_STAGES) & 1));
                tc_fence_after();
                const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * CORR_BN);
                uint32_t v0[32], v1[32];
                const int row0 = r0 + 2 * h, row1 = row0 + 1;
                const int tj0 = row0 - a, tj1 = row1 - a;
                const bool use0 = tj0 >= -CD && tj0 <= CD && row0 < t.H2c && !(p.dbg_skip & 16);     // warp-uniform
                const bool use1 = tj1 >= -CD && tj1 <= CD && row1 < t.H2c && !(p.dbg_skip & 16);
                if (use0) TMEM_LD_32(t_row + (2 * h) * 32, v0);
                if (use1) TMEM_LD_32(t_row + (2 * h + 1) * 32, v1);
                if (use0 || use1) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                tc_fence_before();
                tmem_stage_release(map_to_cta(acce_bar(as), 0));   // this thread is done with the TMEM stage (own CTA = cluster rank 0)
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const bool use = rr ? use1 : use0;
                    if (!use) continue;
                    const int tj = rr ? tj1 : tj0;
                    __syncwarp();
#pragma unroll
                    for (int c = 0; c < 32; ++c) S[c * 34 + lane] = __uint_as_float(rr ? v1[c] : v0[c]);
                    __syncwarp();
                    if (valid) {
                        float* oc = o + (size_t)((tj + CD) * CDS) * plane;
#pragma unroll
                        for (int ti = 0; ti < CDS; ++ti) {
                            const int b2 = lane + ti - CD;
                            float val = (b2 >= 0 && b2 < 32) ? S[b2 * 34 + lane] : 0.f;
                            val *= p.inv;
                            if (p.leaky) val = val > 0.f ? val : 0.1f * val;
                            oc[(size_t)ti * plane] = val;
                        }
                    }
                }
            }
        }
    }
    if (prof && lane == 0 && (warp == 0 || warp >= C3_EPI_WARPS)) {
        const int role = warp == 0 ? 2 : warp - C3_EPI_WARPS;          // 0 = MMA, 1 = producer, 2 = epilogue warp 0
        long long* o = g_corr_prof[blockIdx.x] + role * 4;
        o[0] = pw[0], o[1] = pw[1], o[2] = pw[2], o[3] = clock64() - t_start;
        if (role == 0) g_corr_prof[blockIdx.x][12] = pw[4], g_corr_prof[blockIdx.x][13] = pw[5], g_corr_prof[blockIdx.x][14] = pw[6];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == C3_EPI_WARPS) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

template <bool SPLIT>
static int launch_corr3(const CorrK& k, cudaStream_t s) {
    using G = Corr3Geo<SPLIT>;
    static bool attr_set[MAX_DEVICES];
    if (ensure_dyn_smem(corr_tc3_kernel<SPLIT>, G::SMEM, attr_set, "corr_tc3_kernel")) return 1;
    const int grid = k.num_units < device_sms() ? k.num_units : device_sms();
    corr_tc3_kernel<SPLIT><<<grid, C3_THREADS, G::SMEM, s>>>(k);
    return check_launch("corr_tc3_kernel");
}

// Parity-planar cost volume [B,441,4,H2,W2] fp32 -> channels [0,441) of a split-fp16 NHWC view (FlowNetC's 473-channel merge
// buffer); channels [441, dst.c) are zero-filled.  One thread per (pixel, 8-channel group): 128-byte coalesced reads along x
// within a parity plane, 16-byte stores.
__global__ void __launch_bounds__(256) corr_planar_pack_kernel(const float* __restrict__ src, ft_act d, int H2, int W2) {
    const int groups = d.c / 8;
    const size_t npix = (size_t)d.n * d.h * d.w;
    const size_t cplane = (size_t)4 * H2 * W2;
    __half* dh = reinterpret_cast<__half*>(d.hi);
    __half* dl = reinterpret_cast<__half*>(d.lo);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < npix * groups; i += (size_t)gridDim.x * blockDim.x) {
        // consecutive threads walk the columns of ONE parity class first, so the planar reads coalesce
        const size_t spix = i % npix;
        const int g = (int)(i / npix);
        const int per_img = d.h * d.w;
        const int n = (int)(spix / per_img), r = (int)(spix % per_img);
        const int cls = r / (H2 * W2), rr = r % (H2 * W2);
        const int a = rr / W2, bcol = rr % W2;
        const int y = 2 * a + (cls >> 1), x = 2 * bcol + (cls & 1);
        if (y >= d.h || x >= d.w) continue;
        const float* s = src + (size_t)n * (CDS * CDS) * cplane + ((size_t)cls * H2 + a) * W2 + bcol;
        Half8 h8, l8;
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) {
            const int c0 = g * 8 + 2 * k2;
            const float v0 = c0 < CDS * CDS ? __ldg(s + (size_t)c0 * cplane) : 0.f, v1 = c0 + 1 < CDS * CDS ? __ldg(s + (size_t)(c0 + 1) * cplane) : 0.f;
            __half h0, l0, h1, l1;
            split_f32(v0, h0, l0);
            split_f32(v1, h1, l1);
            h8.v[k2] = __halves2half2(h0, h1), l8.v[k2] = __halves2half2(l0, l1);
        }
        const size_t off = (((size_t)n * d.h + y) * d.w + x) * d.ld + d.coff + g * 8;
        *reinterpret_cast<Half8*>(dh + off) = h8;
        if (dl) *reinterpret_cast<Half8*>(dl + off) = l8;
    }
}

template <bool SPLIT>
static int launch_corr(const CorrK& k, cudaStream_t s) {
    using G = CorrGeo<SPLIT>;
    static bool attr_set[MAX_DEVICES];
    if (ensure_dyn_smem(corr_tc_kernel<SPLIT>, G::SMEM, attr_set, "corr_tc_kernel")) return 1;
    const int grid = k.num_units < device_sms() ? k.num_units : device_sms();
    corr_tc_kernel<SPLIT><<<grid, CORR_THREADS, G::SMEM, s>>>(k);
    return check_launch("corr_tc_kernel");
}

static int check_corr_view(const ft_act& a, const char* name, bool split) {
    FT_REQUIRE(a.hi && (!split || a.lo), "%s: null plane", name);
    FT_REQUIRE(a.c % 8 == 0 && a.ld % 8 == 0 && a.coff % 8 == 0 && a.coff + a.c <= a.ld, "%s: c/ld/coff must be multiples of 8 (c=%d ld=%d coff=%d)", name, a.c, a.ld, a.coff);
    FT_REQUIRE(((uintptr_t)a.hi & 15) == 0 && ((uintptr_t)a.lo & 15) == 0, "%s: planes must be 16-byte aligned", name);
    return 0;
}

}  // namespace ft

using namespace ft;

extern "C" int ft_correlation_tc_supported(int C, int H, int W, int pad, int ksize, int max_disp, int stride1, int stride2) {
    return pad == 20 && ksize == 1 && max_disp == 20 && stride1 == 1 && stride2 == 2 && C % 64 == 0 && C >= 64 && C <= 256 && H >= 1 && W >= 1;
}

static int correlation_act_impl(const ft_act* in1, const ft_act* in2, float* out, int fuse_leaky, int precision, int planar, void* stream);

extern "C" int ft_correlation_act(const ft_act* in1, const ft_act* in2, float* out, int fuse_leaky, int precision, void* stream) {
    return correlation_act_impl(in1, in2, out, fuse_leaky, precision, 0, stream);
}

extern "C" int ft_correlation_act_planar(const ft_act* in1, const ft_act* in2, float* out, int fuse_leaky, int precision, void* stream) {
    return correlation_act_impl(in1, in2, out, fuse_leaky, precision, 1, stream);
}

static int correlation_act_impl(const ft_act* in1, const ft_act* in2, float* out, int fuse_leaky, int precision, int planar, void* stream) {
    FT_REQUIRE(in1 && in2 && out, "correlation_act: null argument");
    FT_REQUIRE(precision == FT_PREC_SPLIT || precision == FT_PREC_FAST, "correlation_act: bad precision %d", precision);
    const bool split = precision == FT_PREC_SPLIT;
    if (int e = check_corr_view(*in1, "correlation_act.in1", split)) return e;
    if (int e = check_corr_view(*in2, "correlation_act.in2", split)) return e;
    FT_REQUIRE(in1->n == in2->n && in1->h == in2->h && in1->w == in2->w && in1->c == in2->c && in1->n > 0, "correlation_act: the two inputs differ in shape");
    FT_REQUIRE(ft_correlation_tc_supported(in1->c, in1->h, in1->w, 20, 1, 20, 1, 2),
               "correlation_act: the tensor-core path needs C in {64,128,192,256} (got C=%d H=%d W=%d)", in1->c, in1->h, in1->w);
    CorrK k = {};
    k.out = out, k.B = in1->n, k.H = in1->h, k.W = in1->w, k.C = in1->c, k.kb = in1->c / CORR_BK;
    k.inv = 1.f / (float)k.C, k.leaky = fuse_leaky != 0, k.planar = planar;
    k.dbg_skip = dev_knob(FT_DBG_SKIP);
    const int H2 = (k.H + 1) / 2;
    k.a_tiles = ceil_div(H2, CORR_ROWS);
    const int W2 = (k.W + 1) / 2;
    k.b_tiles = ceil_div(W2, CORR_COLS), k.ncc = W2 > CORR_COLS ? 2 : 1;
    k.num_tiles = k.B * 4 * k.a_tiles;
    // work units: a tile's chunk list is cut into 3 parts of <= 2 chunks (each part re-loads the tile's A operand through
    // L2): whole tiles (4-6 chunks) leave the last round of the 148 persistent CTAs mostly idle.  Measured on B200 at
    // [B,256,48,64], parts = 1/2/3/4/5/6: B=32 142/143/132/178/133/148 us, B=16 68/74/66/86/64/72, B=8 51/43/41/51/43/43
    // (parts sharing a factor with 148 pin the large parts to the same CTAs under the static round-robin order).
    k.parts = 3;
    if (const int v = dev_knob(FT_DBG_CORR_PARTS)) k.parts = v >= 1 && v <= 6 ? v : 3;
    k.num_units = k.num_tiles * k.parts;
    const int box[4] = {CORR_BK, CORR_COLS, CORR_ROWS, 1};
    for (int cls = 0; cls < 4; ++cls) {
        const int py = cls >> 1, px = cls & 1;
        if (py >= k.H || px >= k.W) {      // degenerate 1-pixel-wide/high image: the class is empty; map it onto class 0 (nothing valid is stored)
            for (int pl = 0; pl < 2; ++pl) k.a_map[cls][pl] = k.a_map[0][pl], k.b_map[cls][pl] = k.b_map[0][pl], k.b2_map[cls][pl] = k.b2_map[0][pl];
            continue;
        }
        for (int op = 0; op < 2; ++op) {
            const ft_act& a = op ? *in2 : *in1;
            const long long dim[4] = {a.c, (k.W - px + 1) / 2, (k.H - py + 1) / 2, k.B};
            const long long str[3] = {2ll * a.ld, 2ll * k.W * a.ld, (long long)k.H * k.W * a.ld};
            const size_t off = (size_t)(py * k.W + px) * a.ld + a.coff;
            CUtensorMap* m = op ? k.b_map[cls] : k.a_map[cls];
            if (int e = make_map(&m[0], (const __half*)a.hi + off, dim, str, box, 128, op ? "corr.in2.hi" : "corr.in1.hi")) return e;
            if (split)
                if (int e = make_map(&m[1], (const __half*)a.lo + off, dim, str, box, 128, op ? "corr.in2.lo" : "corr.in1.lo")) return e;
            if (op) {
                const int box2[4] = {CORR_BK, CORR_COLS, CORR_ROWS / 2, 1};
                if (int e = make_map(&k.b2_map[cls][0], (const __half*)a.hi + off, dim, str, box2, 128, "corr.in2.hi(half)")) return e;
                if (split)
                    if (int e = make_map(&k.b2_map[cls][1], (const __half*)a.lo + off, dim, str, box2, 128, "corr.in2.lo(half)")) return e;
            }
        }
    }
    // 2 = single CTA with resident A (default), 4 = CTA pairs (cta_group::2; the only kernel that tiles maps wider than 64 px),
    // 3 = A operand in TMEM (experimental).  The pair kernel was the default while it was the fastest (105 vs 116 us at
    // [32,256,48,64] on B200 boxes with driver 580.159); on the 580.178 boxes the plain global stores (STG) of a CLUSTER-launched
    // kernel run ~6x slower -- the same pair binary takes 255 us there, 81 us with its stores switched off, while the single-CTA
    // kernel (no cluster) measures 116 us on both (profiles/r02x_corr_cluster_store_regression.txt) -- so the default is the
    // kernel that does not depend on it.  The convolution kernels are not affected: their cluster variants store through TMA.
    int variant = 2;
    if (const int v = dev_knob(FT_DBG_CORR_VARIANT)) variant = v;
    if (k.ncc == 2) variant = 4;                  // frames wider than 64 px: only the CTA-pair kernel tiles the columns
    if (k.planar && variant == 3) variant = 2;    // the TMEM-operand experiment writes NCHW only
    if (variant == 4) return split ? launch_corr_pair<true>(k, (cudaStream_t)stream) : launch_corr_pair<false>(k, (cudaStream_t)stream);
    if (variant == 3) return split ? launch_corr3<true>(k, (cudaStream_t)stream) : launch_corr3<false>(k, (cudaStream_t)stream);
    return split ? launch_corr<true>(k, (cudaStream_t)stream) : launch_corr<false>(k, (cudaStream_t)stream);
}

extern "C" size_t ft_correlation_workspace_bytes(int B, int C, int H, int W) {
    // two split-fp16 NHWC copies of the inputs (hi and lo planes each), 256-byte aligned
    const size_t plane = (((size_t)B * H * W * C * 2) + 255) & ~(size_t)255;
    return 4 * plane;
}

extern "C" int ft_correlation_forward_ws(const float* in1, const float* in2, float* out, int B, int C, int H, int W, int pad, int ksize, int max_disp,
                                         int stride1, int stride2, int fuse_leaky, int precision, void* workspace, size_t workspace_bytes, void* stream) {
    FT_REQUIRE(in1 && in2 && out && B > 0 && C > 0 && H > 0 && W > 0, "correlation_ws: bad arguments");
    if (!ft_correlation_tc_supported(C, H, W, pad, ksize, max_disp, stride1, stride2))
        return ft_correlation_forward(in1, in2, out, B, C, H, W, pad, ksize, max_disp, stride1, stride2, fuse_leaky, stream);
    FT_REQUIRE(workspace && ((uintptr_t)workspace & 255) == 0, "correlation_ws: the workspace must be 256-byte aligned");
    FT_REQUIRE(workspace_bytes >= ft_correlation_workspace_bytes(B, C, H, W), "correlation_ws: workspace too small (%zu < %zu bytes)", workspace_bytes,
               ft_correlation_workspace_bytes(B, C, H, W));
    const size_t plane = ft_correlation_workspace_bytes(B, C, H, W) / 4;
    ft_act a[2];
    for (int i = 0; i < 2; ++i) {
        a[i] = ft_act{(uint8_t*)workspace + (2 * i) * plane, (uint8_t*)workspace + (2 * i + 1) * plane, B, H, W, C, C, 0};
        const long long npix = (long long)H * W;
        corr_relayout_kernel<<<dim3((unsigned)ceil_div((int)npix, 32), C / 64, B), 256, 0, (cudaStream_t)stream>>>(i ? in2 : in1, (__half*)a[i].hi, (__half*)a[i].lo, C, npix);
        if (check_launch("corr_relayout_kernel")) return 1;
    }
    return ft_correlation_act(&a[0], &a[1], out, fuse_leaky, precision, stream);
}

// dev tool: copy the per-CTA wait-cycle counters of the last instrumented corr_tc3_kernel launch (FT_DEV_SKIP bit 64) to the host
extern "C" int ft_dev_corr_profile(long long* host_out /* [148][16] */) {
    return cudaMemcpyFromSymbol(host_out, g_corr_prof, sizeof(long long) * 148 * 16) == cudaSuccess ? 0 : 1;
}

extern "C" int ft_pack_corr_planar(const float* planar, const ft_act* dst, void* stream) {
    FT_REQUIRE(planar && dst && dst->hi, "pack_corr_planar: null pointer");
    FT_REQUIRE(dst->c % 8 == 0 && dst->c >= CDS * CDS && dst->ld % 8 == 0 && dst->coff % 8 == 0 && dst->coff + dst->c <= dst->ld,
               "pack_corr_planar: the destination view needs >= 441 channels in multiples of 8 (c=%d ld=%d coff=%d)", dst->c, dst->ld, dst->coff);
    const int H2 = (dst->h + 1) / 2, W2 = (dst->w + 1) / 2;
    // the index space is [n][4 classes][H2][W2] x groups; for odd H or W the absent pixels of a class are skipped
    ft_act d = *dst;
    const size_t total = (size_t)d.n * 4 * H2 * W2 * (d.c / 8);
    FT_REQUIRE((size_t)4 * H2 * W2 == (size_t)d.h * d.w, "pack_corr_planar: H and W must be even (got %d x %d)", d.h, d.w);
    const size_t cap = (size_t)device_sms() * 32;
    const int grid = (int)((total + 255) / 256 < cap ? (total + 255) / 256 : cap);
    corr_planar_pack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(planar, d, H2, W2);
    return check_launch("corr_planar_pack_kernel");
}
