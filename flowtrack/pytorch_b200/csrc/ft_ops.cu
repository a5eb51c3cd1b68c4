// ft_ops.cu -- the HBM-bound kernels of the path: the three native FlowNet2 ops (Correlation,
// Resample2d, ChannelNorm), layout packing, max-pool, x4 upsampling, FlowNet2 stacking glue and
// key-point decode / box propagation.  sm_100a only.  See include/flowtrack_b200.h for the ABI and
// the reference lines each entry point replaces.
#include <math.h>
#include <stdarg.h>

#include <atomic>

#include "ft_common.cuh"

namespace ft {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches += n; }

int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
    return dev < MAX_DEVICES ? dev : MAX_DEVICES - 1;
}
int device_sms() {
    static std::atomic<int> sms[MAX_DEVICES];
    const int dev = current_device();
    int v = sms[dev].load(std::memory_order_relaxed);
    if (v == 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;   // B200
        sms[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}
static std::atomic<int> g_knobs[FT_DBG_COUNT];
int dev_knob(int which) { return (which >= 0 && which < FT_DBG_COUNT) ? g_knobs[which].load(std::memory_order_relaxed) : 0; }
int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return 1;
    }
    count_launch();
    return 0;
}

// =====================================================================================================
// Correlation (A5).  out[n, (tj+d)*(2d+1)+(ti+d), y, x] = 1/(k*k*C) * sum_{j,i,c} in1[n,c,y1+j,x1+i] * in2[n,c,y2+j,x2+i]
// with (y1,x1) = (y,x)*stride1 + max_disp + krad - pad in un-padded coordinates and (y2,x2) = (y1,x1) + (tj,ti)*stride2;
// reads outside the image contribute zero (the reference materialises a zero-padded copy instead).
// =====================================================================================================

// Generic parameters: lane = output pixel along x (coalesced NCHW reads), each thread owns one
// displacement row tj and sweeps ti in registers; channels are summed per thread (no cross-lane step).
__global__ void corr_generic_kernel(const float* __restrict__ in1, const float* __restrict__ in2, float* __restrict__ out,
                                    int C, int H, int W, int outH, int outW, int pad, int krad, int max_disp, int s1, int s2,
                                    int drad, int leaky) {
    const int dsize = 2 * drad + 1;
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y;
    const int n = blockIdx.z;
    if (x >= outW) return;
    const int y1 = y * s1 + max_disp + krad - pad;
    const int x1 = x * s1 + max_disp + krad - pad;
    const float inv = 1.f / (float)((2 * krad + 1) * (2 * krad + 1) * C);
    const size_t img = (size_t)C * H * W;
    const float* a = in1 + n * img;
    const float* b = in2 + n * img;
    for (int t = threadIdx.y; t < dsize * dsize; t += blockDim.y) {
        const int tj = t / dsize - drad, ti = t % dsize - drad;
        float acc = 0.f;
        for (int j = -krad; j <= krad; ++j)
            for (int i = -krad; i <= krad; ++i) {
                const int ya = y1 + j, xa = x1 + i, yb = ya + tj * s2, xb = xa + ti * s2;
                if (ya < 0 || ya >= H || xa < 0 || xa >= W || yb < 0 || yb >= H || xb < 0 || xb >= W) continue;
                const float* pa = a + (size_t)ya * W + xa;
                const float* pb = b + (size_t)yb * W + xb;
                for (int c = 0; c < C; ++c) acc = fmaf(pa[(size_t)c * H * W], pb[(size_t)c * H * W], acc);
            }
        float v = acc * inv;
        if (leaky) v = v > 0.f ? v : 0.1f * v;
        out[(((size_t)n * dsize * dsize + t) * outH + y) * outW + x] = v;
    }
}

// FlowNetC fast path: ksize 1, stride1 1, stride2 2, displacement radius 10 (x2 px), pad = 20.
// Because stride2 = 2, a pixel only ever meets pixels of its own column parity.  A block owns
// CORR_R output rows x 64 columns x CORR_TJ displacement rows; per channel chunk it stages
//   in1  [CC][CORR_R][2 parities][32]
//   in2  [CC][CORR_R + 2*(CORR_TJ-1)][2 parities][52]   (20 px halo each side, zero outside)
// in shared memory (parity-de-interleaved so that every operand fetch is an aligned LDS.128), and each
// thread accumulates 4 same-parity pixels x 21 horizontal displacements = 84 sums in registers.
#define CORR_R 8
#define CORR_TJ 3
#define CORR_CC 4
#define CORR_ROWS2 (CORR_R + 2 * (CORR_TJ - 1))
#define CORR_THREADS (CORR_R * CORR_TJ * 16)

__global__ void __launch_bounds__(CORR_THREADS) corr_fast_kernel(const float* __restrict__ in1, const float* __restrict__ in2,
                                                                 float* __restrict__ out, int C, int H, int W, int leaky) {
    __shared__ __align__(16) float s1[CORR_CC][CORR_R][2][32];
    __shared__ __align__(16) float s2[CORR_CC][CORR_ROWS2][2][52];
    const int tid = threadIdx.x;
    const int xg = tid & 15;                 // 16 groups of 4 same-parity pixels per 64-px row segment
    const int par = xg & 1, xq = xg >> 1;    // pixel k of the group sits at column x0 + 2*(4*xq + k) + par
    const int tjl = (tid >> 4) % CORR_TJ;
    const int r = tid / (16 * CORR_TJ);
    const int x0 = blockIdx.x * 64;
    const int ntj_groups = 21 / CORR_TJ;
    const int y0 = (blockIdx.y / ntj_groups) * CORR_R;
    const int tj0 = (blockIdx.y % ntj_groups) * CORR_TJ - 10;
    const int n = blockIdx.z;
    const size_t plane = (size_t)H * W;
    const float* a = in1 + (size_t)n * C * plane;
    const float* b = in2 + (size_t)n * C * plane;

    float acc[4][21];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int t = 0; t < 21; ++t) acc[k][t] = 0.f;

    for (int c0 = 0; c0 < C; c0 += CORR_CC) {
        __syncthreads();
        for (int e = tid; e < CORR_CC * CORR_R * 64; e += CORR_THREADS) {
            const int xx = e & 63, rr = (e >> 6) % CORR_R, cc = e / (64 * CORR_R);
            const int gy = y0 + rr, gx = x0 + xx, gc = c0 + cc;
            float v = 0.f;
            if (gc < C && gy < H && gx < W) v = __ldg(a + gc * plane + (size_t)gy * W + gx);
            s1[cc][rr][xx & 1][xx >> 1] = v;
        }
        for (int e = tid; e < CORR_CC * CORR_ROWS2 * 104; e += CORR_THREADS) {
            const int xx = e % 104, rr = (e / 104) % CORR_ROWS2, cc = e / (104 * CORR_ROWS2);
            const int gy = y0 + 2 * tj0 + rr, gx = x0 - 20 + xx, gc = c0 + cc;
            float v = 0.f;
            if (gc < C && gy >= 0 && gy < H && gx >= 0 && gx < W) v = __ldg(b + gc * plane + (size_t)gy * W + gx);
            s2[cc][rr][xx & 1][xx >> 1] = v;   // (x0-20) is even, so column parity is preserved
        }
        __syncthreads();
#pragma unroll 2
        for (int cc = 0; cc < CORR_CC; ++cc) {
            const float4 p = *reinterpret_cast<const float4*>(&s1[cc][r][par][4 * xq]);
            const float pv[4] = {p.x, p.y, p.z, p.w};
            float q[24];
#pragma unroll
            for (int v4 = 0; v4 < 6; ++v4) {
                const float4 t = *reinterpret_cast<const float4*>(&s2[cc][r + 2 * tjl][par][4 * xq + 4 * v4]);
                q[4 * v4 + 0] = t.x, q[4 * v4 + 1] = t.y, q[4 * v4 + 2] = t.z, q[4 * v4 + 3] = t.w;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int t = 0; t < 21; ++t) acc[k][t] = fmaf(pv[k], q[k + t], acc[k][t]);
        }
    }
    const int y = y0 + r;
    if (y >= H) return;
    const float inv = 1.f / (float)C;
    const int tj = tj0 + tjl;
#pragma unroll
    for (int t = 0; t < 21; ++t) {
        const int ch = (tj + 10) * 21 + t;
        float* o = out + (((size_t)n * 441 + ch) * H + y) * W;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int x = x0 + 2 * (4 * xq + k) + par;
            if (x < W) {
                float v = acc[k][t] * inv;
                if (leaky) v = v > 0.f ? v : 0.1f * v;
                o[x] = v;
            }
        }
    }
}

// =====================================================================================================
// Resample2d (A7) and ChannelNorm (A8)
// =====================================================================================================
struct Bilin {
    int xL, xR, yT, yB;
    float w00, w01, w10, w11;
};
__device__ __forceinline__ Bilin bilin_setup(float xf, float yf, int W, int H) {
    Bilin s;
    const float fx = floorf(xf), fy = floorf(yf);
    const float alpha = xf - fx, beta = yf - fy;
    s.xL = max(min((int)fx, W - 1), 0);
    s.xR = max(min((int)(fx + 1.f), W - 1), 0);
    s.yT = max(min((int)fy, H - 1), 0);
    s.yB = max(min((int)(fy + 1.f), H - 1), 0);
    s.w00 = (1.f - alpha) * (1.f - beta);
    s.w01 = alpha * (1.f - beta);
    s.w10 = (1.f - alpha) * beta;
    s.w11 = alpha * beta;
    return s;
}
__device__ __forceinline__ float bilin_sample(const float* __restrict__ p, const Bilin& s, int W) {
    float v = s.w00 * __ldg(p + (size_t)s.yT * W + s.xL);
    v += s.w01 * __ldg(p + (size_t)s.yT * W + s.xR);
    v += s.w10 * __ldg(p + (size_t)s.yB * W + s.xL);
    v += s.w11 * __ldg(p + (size_t)s.yB * W + s.xR);
    return v;
}

// Per-pixel kernels: x from blockIdx.x / threadIdx.x, rows from blockIdx.y, images from blockIdx.z (the latter two grid-strided past
// 65535) -- no 64-bit division per element (the grid-stride-over-a-linear-index form spent most of its instructions on them).
#define FT_FOR_PIXELS(B_, H_, W_, bvar_, yvar_, xvar_)                         \
    const int xvar_ = blockIdx.x * blockDim.x + threadIdx.x;                   \
    if (xvar_ < (W_))                                                          \
        for (int bvar_ = blockIdx.z; bvar_ < (B_); bvar_ += gridDim.z)         \
            for (int yvar_ = blockIdx.y; yvar_ < (H_); yvar_ += gridDim.y)
constexpr int PIX_THREADS = 128;

__global__ void __launch_bounds__(PIX_THREADS) resample2d_kernel(const float* __restrict__ img, const float* __restrict__ flow, float* __restrict__ out,
                                                                 int B, int C, int H, int W) {
    const size_t plane = (size_t)H * W;
    FT_FOR_PIXELS(B, H, W, b, y, x) {
        const size_t p = (size_t)y * W + x;
        const float dx = flow[((size_t)b * 2) * plane + p];
        const float dy = flow[((size_t)b * 2 + 1) * plane + p];
        const Bilin s = bilin_setup((float)x + dx, (float)y + dy, W, H);
        for (int c = 0; c < C; ++c) out[((size_t)b * C + c) * plane + p] = bilin_sample(img + ((size_t)b * C + c) * plane, s, W);
    }
}

__global__ void __launch_bounds__(PIX_THREADS) channelnorm_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int C, int H, int W) {
    const size_t plane = (size_t)H * W;
    FT_FOR_PIXELS(B, H, W, b, y, px) {
        const size_t p = (size_t)y * W + px;
        float r = 0.f;
        for (int c = 0; c < C; ++c) {
            const float v = x[((size_t)b * C + c) * plane + p];
            r = fmaf(v, v, r);
        }
        out[(size_t)b * plane + p] = sqrtf(r);
    }
}

// =====================================================================================================
// Layout packing / pooling / upsampling
// =====================================================================================================
__device__ __forceinline__ void store8(const ft_act& d, size_t pix, int c8, const float v[8]) {
    Half8 hi, lo;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        __half h0, l0, h1, l1;
        split_f32(v[2 * k], h0, l0);
        split_f32(v[2 * k + 1], h1, l1);
        hi.v[k] = __halves2half2(h0, h1);
        lo.v[k] = __halves2half2(l0, l1);
    }
    const size_t off = pix * d.ld + d.coff + c8;
    *reinterpret_cast<Half8*>(reinterpret_cast<__half*>(d.hi) + off) = hi;
    if (d.lo) *reinterpret_cast<Half8*>(reinterpret_cast<__half*>(d.lo) + off) = lo;
}

__global__ void pack_image_kernel(const float* __restrict__ src, long long sb, long long sc, long long sh, long long sw,
                                  int C, float mul, ft_act d, int W, int xl) {
    // d.w is the physical row pitch of the destination (image width W plus zero columns left/right); xl = left padding
    const size_t npix = (size_t)d.n * d.h * W;
    const int groups = d.c / 8;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < npix * groups; i += (size_t)gridDim.x * blockDim.x) {
        // consecutive threads walk x first so the strided NCHW reads coalesce per channel
        const size_t spix = i % npix;
        const int g = (int)(i / npix);
        const int x = (int)(spix % W), y = (int)((spix / W) % d.h), b = (int)(spix / ((size_t)W * d.h));
        const size_t pix = ((size_t)b * d.h + y) * d.w + x + xl;
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int c = g * 8 + k;
            v[k] = c < C ? mul * __ldg(src + b * sb + c * sc + y * sh + x * sw) : 0.f;
        }
        store8(d, pix, g * 8, v);
    }
}

// space-to-depth packing of a 3-channel image (see ft_pack_image_s2d): one thread per s2d pixel = 16 channels
__global__ void __launch_bounds__(PIX_THREADS) pack_image_s2d_kernel(const float* __restrict__ src, int B, int H, int W, ft_act d, int xl) {
    const int H2 = H >> 1, W2 = W >> 1;
    const size_t plane = (size_t)H * W;
    FT_FOR_PIXELS(B, H2, W2, b, Y, X) {
        const float* s = src + (size_t)b * 3 * plane + (size_t)(2 * Y) * W + 2 * X;
        float v[16];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float2 r0 = __ldg(reinterpret_cast<const float2*>(s + c * plane)), r1 = __ldg(reinterpret_cast<const float2*>(s + c * plane + W));
            v[0 + c] = r0.x, v[4 + c] = r0.y, v[8 + c] = r1.x, v[12 + c] = r1.y;
        }
        v[3] = v[7] = v[11] = v[15] = 0.f;
        const size_t pix = ((size_t)b * H2 + Y) * d.w + X + xl;
        store8(d, pix, 0, v);
        store8(d, pix, 8, v + 8);
    }
}

// grid: x = (output column, 8-channel group) pairs, y = output rows, z = images -- no 64-bit index arithmetic per element (the
// grid-stride form spent most of its instructions on size_t divisions)
__global__ void __launch_bounds__(256) maxpool3x3s2_kernel(ft_act in, ft_act out) {
    const int groups = out.c / 8;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int ox = t / groups, g = t - ox * groups;
    if (ox >= out.w) return;
    const __half* ih = reinterpret_cast<const __half*>(in.hi);
    const __half* il = reinterpret_cast<const __half*>(in.lo);
    for (int b = blockIdx.z; b < out.n; b += gridDim.z)
        for (int oy = blockIdx.y; oy < out.h; oy += gridDim.y) {
            float best[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) best[k] = -INFINITY;
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) {
                    const int iy = oy * 2 + dy, ix = ox * 2 + dx;
                    if (iy < 0 || iy >= in.h || ix < 0 || ix >= in.w) continue;
                    const size_t off = (((size_t)b * in.h + iy) * in.w + ix) * in.ld + in.coff + g * 8;
                    const Half8 h = *reinterpret_cast<const Half8*>(ih + off);
                    Half8 l;
                    if (il) l = *reinterpret_cast<const Half8*>(il + off);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        float2 f = __half22float2(h.v[k]);
                        if (il) {
                            const float2 fl = __half22float2(l.v[k]);
                            f.x += fl.x, f.y += fl.y;
                        }
                        best[2 * k] = fmaxf(best[2 * k], f.x);
                        best[2 * k + 1] = fmaxf(best[2 * k + 1], f.y);
                    }
                }
            store8(out, ((size_t)b * out.h + oy) * out.w + ox, g * 8, best);
        }
}

#define PAIR_MEAN_CHUNKS 64
// mean over both frames and all pixels per (b, colour) (models.py:112), in two deterministic steps:
// blockIdx.y-th chunk partial sums in double, then one warp per (b, colour) folds the partials.
__global__ void pair_mean_partial_kernel(const float* __restrict__ pair, int HW2, double* __restrict__ partial) {
    const float* p = pair + (size_t)blockIdx.x * HW2;
    const int per = (HW2 + PAIR_MEAN_CHUNKS - 1) / PAIR_MEAN_CHUNKS;
    const int lo = blockIdx.y * per, hi = min(HW2, lo + per);
    double s = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int i = lo + threadIdx.x;
    for (; i + 3 * (int)blockDim.x < hi; i += 4 * blockDim.x) {
        s += (double)p[i], s1 += (double)p[i + blockDim.x], s2 += (double)p[i + 2 * blockDim.x], s3 += (double)p[i + 3 * blockDim.x];
    }
    for (; i < hi; i += blockDim.x) s += (double)p[i];
    s += s1 + s2 + s3;
    __shared__ double sh[32];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) partial[blockIdx.x * PAIR_MEAN_CHUNKS + blockIdx.y] = s;
    }
}
__global__ void pair_mean_final_kernel(const double* __restrict__ partial, int HW2, float* __restrict__ mean) {
    double s = 0.0;
    for (int i = threadIdx.x; i < PAIR_MEAN_CHUNKS; i += 32) s += partial[blockIdx.x * PAIR_MEAN_CHUNKS + i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) mean[blockIdx.x] = (float)(s / (double)HW2);
}

__global__ void __launch_bounds__(PIX_THREADS) pack_pair_kernel(const float* __restrict__ pair, const float* __restrict__ mean, int B, int H, int W,
                                                                float rgb_max, float* __restrict__ x_f32, ft_act d, int has_dst, int xl) {
    const size_t plane = (size_t)H * W;
    FT_FOR_PIXELS(B, H, W, b, y, x) {
        const size_t p = (size_t)y * W + x;
        float v[8];
#pragma unroll
        for (int f = 0; f < 2; ++f)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                // pair is [B,3,2,H,W]; x channel = f*3 + c  (models.py:116-119)
                const float t = (pair[(((size_t)b * 3 + c) * 2 + f) * plane + p] - mean[b * 3 + c]) / rgb_max;
                v[f * 3 + c] = t;
                if (x_f32) x_f32[((size_t)b * 6 + f * 3 + c) * plane + p] = t;
            }
        v[6] = v[7] = 0.f;
        if (has_dst) {
            const size_t dpix = ((size_t)b * H + y) * d.w + x + xl;       // d.w = physical row pitch (W + zero columns)
            store8(d, dpix, 0, v);
            const float z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int g = 1; g < d.c / 8; ++g) store8(d, dpix, g * 8, z);
        }
    }
}

// ---- uint8 frame pairs (the reference's cv2 frames, net_utils.py:83-87): frames [B,2,H,W,3] uint8, channel order BGR (cv2) or RGB ----
// per (b, memory channel) sums over both frames: exact integer arithmetic, one atomicAdd per block and channel
__global__ void __launch_bounds__(256) pair_u8_sum_kernel(const uint8_t* __restrict__ frames, size_t bytes_per_pair, unsigned long long* __restrict__ sums) {
    const int b = blockIdx.y;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(frames + (size_t)b * bytes_per_pair);
    const uint32_t groups = (uint32_t)(bytes_per_pair / 12);        // 12 bytes = 4 pixels: the channel of every byte lane is fixed per word
    uint32_t s0 = 0, s1 = 0, s2 = 0;
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += gridDim.x * blockDim.x) {
        const uint32_t a = __ldg(w + 3 * g), c = __ldg(w + 3 * g + 1), d = __ldg(w + 3 * g + 2);
        // bytes: a = c0 c1 c2 c0 | c = c1 c2 c0 c1 | d = c2 c0 c1 c2  -> byte-select dot products
        s0 = __dp4a(a, 0x01000001u, s0), s0 = __dp4a(c, 0x00010000u, s0), s0 = __dp4a(d, 0x00000100u, s0);
        s1 = __dp4a(a, 0x00000100u, s1), s1 = __dp4a(c, 0x01000001u, s1), s1 = __dp4a(d, 0x00010000u, s1);
        s2 = __dp4a(a, 0x00010000u, s2), s2 = __dp4a(c, 0x00000100u, s2), s2 = __dp4a(d, 0x01000001u, s2);
    }
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o), s1 += __shfl_xor_sync(0xffffffffu, s1, o), s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    // one atomicAdd per block and channel (integers: the result does not depend on the order)
    __shared__ uint32_t sh[3][8];
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) sh[0][warp] = s0, sh[1][warp] = s1, sh[2][warp] = s2;
    __syncthreads();
    if (threadIdx.x < 3) {
        unsigned long long t = 0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += sh[threadIdx.x][k];
        atomicAdd(sums + b * 3 + threadIdx.x, t);
    }
}

// one thread per pixel (a warp's two 16-byte stores per plane are 512 contiguous bytes; a 4-pixels-per-thread variant with wider
// loads measured 2.4x slower on its strided stores); grid: x = columns, y = rows, z = pairs
__global__ void __launch_bounds__(256) pack_pair_u8_kernel(const uint8_t* __restrict__ frames, const unsigned long long* __restrict__ sums, int B, int H, int W, int bgr,
                                                           float rgb_max, float* __restrict__ x_f32, ft_act d, int has_dst, int xl) {
    const size_t plane = (size_t)H * W;
    const double inv_n = 1.0 / (2.0 * (double)plane);
    FT_FOR_PIXELS(B, H, W, b, y, x) {
        const size_t p = (size_t)y * W + x;
        float v[8];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int mc = bgr ? 2 - c : c;                                       // memory channel holding colour c of the RGB flow input
            const float mean = (float)((double)sums[b * 3 + mc] * inv_n);         // models.py:112 (per sample and colour, both frames)
#pragma unroll
            for (int f = 0; f < 2; ++f) {
                const float t = ((float)__ldg(frames + (((size_t)b * 2 + f) * plane + p) * 3 + mc) - mean) / rgb_max;      // x channel = f*3 + c (models.py:116-119)
                v[f * 3 + c] = t;
                if (x_f32) x_f32[((size_t)b * 6 + f * 3 + c) * plane + p] = t;
            }
        }
        v[6] = v[7] = 0.f;
        if (has_dst) {
            const size_t dpix = ((size_t)b * H + y) * d.w + x + xl;               // d.w = physical row pitch (W + zero columns)
            store8(d, dpix, 0, v);
            const float z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int g = 1; g < d.c / 8; ++g) store8(d, dpix, g * 8, z);
        }
    }
}

// ---------------- upsample4x ----------------
// one thread per SOURCE column: the four output columns 4x .. 4x+3 of one output row (a 16-byte store when dst allows it);
// grid: x = source columns, y = output rows, z = planes.  Same arithmetic per element as a one-thread-per-output kernel.
template <int MODE>
__global__ void __launch_bounds__(128) upsample4x_kernel(const float* __restrict__ src, float* __restrict__ dst, int BC, int h, int w, float mul, int vec) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= w) return;
    const int H = 4 * h, W = 4 * w;
    for (int bc = blockIdx.z; bc < BC; bc += gridDim.z) {
        const float* p = src + (size_t)bc * h * w;
        for (int Y = blockIdx.y; Y < H; Y += gridDim.y) {
            float o[4];
            if (MODE == 1) {
                o[0] = o[1] = o[2] = o[3] = p[(Y >> 2) * w + x] * mul;
            } else {
                // align_corners=False: src = (dst + 0.5)/4 - 0.5, clamped below at 0 (ATen area_pixel_compute_source_index)
                const float sy = fmaxf((Y + 0.5f) * 0.25f - 0.5f, 0.f);
                const int y0 = (int)sy, y1 = min(y0 + 1, h - 1);
                const float ly = sy - y0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float sx = fmaxf((4 * x + k + 0.5f) * 0.25f - 0.5f, 0.f);
                    const int x0 = (int)sx, x1 = min(x0 + 1, w - 1);
                    const float lx = sx - x0;
                    const float v = (1.f - ly) * ((1.f - lx) * p[y0 * w + x0] + lx * p[y0 * w + x1]) + ly * ((1.f - lx) * p[y1 * w + x0] + lx * p[y1 * w + x1]);
                    o[k] = v * mul;
                }
            }
            float* q = dst + ((size_t)bc * H + Y) * W + 4 * x;
            if (vec) {
                *reinterpret_cast<float4*>(q) = make_float4(o[0], o[1], o[2], o[3]);
            } else {
                q[0] = o[0], q[1] = o[1], q[2] = o[2], q[3] = o[3];
            }
        }
    }
}

__global__ void __launch_bounds__(PIX_THREADS) warp_concat_kernel(const float* __restrict__ x6, const float* __restrict__ flow, float div_flow, int B, int H, int W, ft_act d) {
    const size_t plane = (size_t)H * W;
    FT_FOR_PIXELS(B, H, W, b, y, x) {
        const size_t p = (size_t)y * W + x;
        const float u = flow[((size_t)b * 2) * plane + p], v = flow[((size_t)b * 2 + 1) * plane + p];
        const Bilin s = bilin_setup((float)x + u, (float)y + v, W, H);
        float o[16];
        float e = 0.f;
#pragma unroll
        for (int c = 0; c < 6; ++c) o[c] = x6[((size_t)b * 6 + c) * plane + p];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            o[6 + c] = bilin_sample(x6 + ((size_t)b * 6 + 3 + c) * plane, s, W);
            const float df = o[c] - o[6 + c];
            e = fmaf(df, df, e);
        }
        o[9] = u / div_flow, o[10] = v / div_flow, o[11] = sqrtf(e);
        o[12] = o[13] = o[14] = o[15] = 0.f;
        store8(d, (size_t)b * plane + p, 0, o);
        store8(d, (size_t)b * plane + p, 8, o + 8);
    }
}

__global__ void __launch_bounds__(PIX_THREADS) fusion_concat_kernel(const float* __restrict__ x6, const float* __restrict__ fsd, const float* __restrict__ fs2,
                                                                    int B, int H, int W, ft_act d) {
    const size_t plane = (size_t)H * W;
    FT_FOR_PIXELS(B, H, W, b, y, x) {
        const size_t p = (size_t)y * W + x;
        float o[16];
        float err[2];
        const float* fl[2] = {fsd, fs2};
#pragma unroll
        for (int c = 0; c < 3; ++c) o[c] = x6[((size_t)b * 6 + c) * plane + p];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const float u = fl[k][((size_t)b * 2) * plane + p], v = fl[k][((size_t)b * 2 + 1) * plane + p];
            o[3 + 2 * k] = u, o[4 + 2 * k] = v;
            o[7 + k] = sqrtf(fmaf(u, u, v * v));
            const Bilin s = bilin_setup((float)x + u, (float)y + v, W, H);
            float e = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float df = o[c] - bilin_sample(x6 + ((size_t)b * 6 + 3 + c) * plane, s, W);
                e = fmaf(df, df, e);
            }
            err[k] = sqrtf(e);
        }
        o[9] = err[0], o[10] = err[1];
        o[11] = o[12] = o[13] = o[14] = o[15] = 0.f;
        store8(d, (size_t)b * plane + p, 0, o);
        store8(d, (size_t)b * plane + p, 8, o + 8);
    }
}

// =====================================================================================================
// Flow prediction head.  predict_flow* is a 3x3 convolution with TWO output channels on a wide concat buffer
// (submodules.py:20-21); as an implicit GEMM its A operand would be re-read nine times for two columns of
// output.  The path instead runs it as a 1x1 convolution with 18 outputs on the tensor cores
//   part[n, t*2 + c, y, x] = sum_ci W[c, ci, r, s] * in[n, ci, y, x],  t = r*3 + s
// (the input is read once) and this kernel finishes the convolution by the shifted sum
//   flow[n, c, y, x] = bias[c] + sum_t part[n, t*2 + c, y + r - 1, x + s - 1]   (zero outside the map).
// =====================================================================================================
// grid: x = columns, y = rows, z = images
__global__ void __launch_bounds__(128) flow_head_kernel(const float* __restrict__ part, const float* __restrict__ bias, float* __restrict__ flow, int B, int H, int W) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    const size_t plane = (size_t)H * W;
    for (int b = blockIdx.z; b < B; b += gridDim.z)
        for (int y = blockIdx.y; y < H; y += gridDim.y) {
            float u = bias ? bias[0] : 0.f, v = bias ? bias[1] : 0.f;
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int s = 0; s < 3; ++s) {
                    const int yy = y + r - 1, xx = x + s - 1;
                    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                    const float* q = part + (((size_t)b * 18 + (r * 3 + s) * 2) * H + yy) * W + xx;
                    u += q[0];
                    v += q[plane];
                }
            const size_t p = (size_t)y * W + x;
            flow[((size_t)b * 2) * plane + p] = u;
            flow[((size_t)b * 2 + 1) * plane + p] = v;
        }
}

// upsampled_flow*_to_*: ConvTranspose2d(2, 2, 4, 2, 1) of the fp32 flow (FlowNetS.py:41-44), written as split-fp16
// NHWC channels [0,2) (+6 zeros) of a slice of the next level's concat buffer.  w is the module's [2(in)][2(out)][4][4].
// grid: x = output columns, y = output rows, z = images
__global__ void __launch_bounds__(128) flow_upsample_kernel(const float* __restrict__ flow, const float* __restrict__ w, const float* __restrict__ bias,
                                                            int B, int h, int wd, ft_act d) {
    __shared__ float sw[64];
    if (threadIdx.x < 64) sw[threadIdx.x] = w[threadIdx.x];
    __syncthreads();
    const int Ho = 2 * h, Wo = 2 * wd;
    const int ox = blockIdx.x * blockDim.x + threadIdx.x;
    if (ox >= Wo) return;
    const size_t plane = (size_t)h * wd;
    for (int b = blockIdx.z; b < B; b += gridDim.z)
        for (int oy = blockIdx.y; oy < Ho; oy += gridDim.y) {
            float o[8] = {bias ? bias[0] : 0.f, bias ? bias[1] : 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            // oy = 2*iy - 1 + ky  ->  ky = oy + 1 - 2*iy in [0,4)
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                const int iy = (oy + 1) / 2 - a, ky = oy + 1 - 2 * iy;
                if (iy < 0 || iy >= h) continue;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int ix = (ox + 1) / 2 - c, kx = ox + 1 - 2 * ix;
                    if (ix < 0 || ix >= wd) continue;
                    const float f0 = flow[((size_t)b * 2) * plane + (size_t)iy * wd + ix], f1 = flow[((size_t)b * 2 + 1) * plane + (size_t)iy * wd + ix];
                    o[0] = fmaf(f0, sw[(0 * 2 + 0) * 16 + ky * 4 + kx], o[0]);
                    o[0] = fmaf(f1, sw[(1 * 2 + 0) * 16 + ky * 4 + kx], o[0]);
                    o[1] = fmaf(f0, sw[(0 * 2 + 1) * 16 + ky * 4 + kx], o[1]);
                    o[1] = fmaf(f1, sw[(1 * 2 + 1) * 16 + ky * 4 + kx], o[1]);
                }
            }
            store8(d, ((size_t)b * Ho + oy) * Wo + ox, 0, o);
        }
}

// =====================================================================================================
// Decode (A9) and box propagation (A10)
// =====================================================================================================
__global__ void decode_kernel(const float* __restrict__ hm, int C, int h, int w, const float* __restrict__ center,
                              const float* __restrict__ scale, int adjust, float threshold, const int32_t* __restrict__ valid,
                              float* __restrict__ kps) {
    const int nc = blockIdx.x;  // n*C + c
    const int n = nc / C;
    if (valid && !valid[n]) {   // an empty person slot of the tracking pipeline: no key-points
        if (threadIdx.x < 3) kps[(size_t)nc * 3 + threadIdx.x] = 0.f;
        return;
    }
    const float* p = hm + (size_t)nc * h * w;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < h * w; i += blockDim.x) {
        const float v = p[i];
        if (v > best) best = v, bi = i;   // ascending i per thread: keeps the first maximum
    }
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) best = ov, bi = oi;
    }
    __shared__ float sv[32];
    __shared__ int si[32];
    if ((threadIdx.x & 31) == 0) sv[threadIdx.x >> 5] = best, si[threadIdx.x >> 5] = bi;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k)
            if (sv[k] > best || (sv[k] == best && si[k] < bi)) best = sv[k], bi = si[k];
        float cx = (float)(bi % w), cy = (float)(bi / w);
        if (!(best > threshold)) cx = 0.f, cy = 0.f;                 // evaluation.py:18-19
        if (adjust) {
            const int x = (int)cx, y = (int)cy;
            if (x > 0 && x < w - 1 && y > 0 && y < h - 1) {          // evaluation.py:31-33
                const float dx = p[y * w + x + 1] - p[y * w + x - 1], dy = p[(y + 1) * w + x] - p[(y - 1) * w + x];
                cx += dx > 0.f ? 0.25f : (dx < 0.f ? -0.25f : 0.f);
                cy += dy > 0.f ? 0.25f : (dy < 0.f ? -0.25f : 0.f);
            }
        }
        // inverse of T = [[s,0,-s*cx0+w/2],[0,s,-s*cy0+h/2]], s = h/scale (transforms.py:180-184), in double like numpy
        const double s = (double)h / (double)scale[n];
        const double X = ((double)cx - (-s * (double)center[2 * n] + 0.5 * w)) / s;
        const double Y = ((double)cy - (-s * (double)center[2 * n + 1] + 0.5 * h)) / s;
        kps[(size_t)nc * 3 + 0] = (float)X;
        kps[(size_t)nc * 3 + 1] = (float)Y;
        kps[(size_t)nc * 3 + 2] = best;
    }
}

__global__ void boxprop_kernel(const float* __restrict__ kps, const int32_t* __restrict__ frame_of, int N, int K,
                               const float* __restrict__ flow, int B, int H, int W, float* __restrict__ boxes) {
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (n >= N) return;
    const int b = frame_of ? frame_of[n] : 0;
    const float big = (float)max(H, W);
    float mnx = INFINITY, mny = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
    for (int k = lane; k < K; k += 32) {
        const float x = kps[((size_t)n * K + k) * 3], y = kps[((size_t)n * K + k) * 3 + 1], sc = kps[((size_t)n * K + k) * 3 + 2];
        const int px = min(max((int)x, 0), W - 1), py = min(max((int)y, 0), H - 1);   // trunc like astype(int), then clamp
        const float sx = x + flow[(((size_t)b * 2) * H + py) * W + px];
        const float sy = y + flow[(((size_t)b * 2 + 1) * H + py) * W + px];
        const bool m = sc > 0.f;
        // flow_utils.py:28-29: min over (shift + (1-mask)*max(H,W)), max over (shift*mask)
        mnx = fminf(mnx, m ? sx : sx + big), mny = fminf(mny, m ? sy : sy + big);
        mxx = fmaxf(mxx, m ? sx : 0.f), mxy = fmaxf(mxy, m ? sy : 0.f);
    }
    for (int o = 16; o > 0; o >>= 1) {
        mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
        mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
    }
    if (lane == 0) {
        const float ex = (mxx - mnx) * 0.15f / 2.f, ey = (mxy - mny) * 0.15f / 2.f;
        boxes[n * 4 + 0] = fmaxf(mnx - ex, 0.f);
        boxes[n * 4 + 1] = fmaxf(mny - ey, 0.f);
        boxes[n * 4 + 2] = fminf(mxx + ex, (float)(W - 1));
        boxes[n * 4 + 3] = fminf(mxy + ey, (float)(H - 1));
    }
}

static int grid_for(size_t total, int threads) {
    size_t g = (total + threads - 1) / threads;
    const size_t cap = (size_t)device_sms() * 16;   // a few resident CTAs per SM, grid-stride beyond that
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

static int check_act(const ft_act* a, const char* name, bool need_lo) {
    FT_REQUIRE(a && a->hi, "%s: null activation view", name);
    FT_REQUIRE(!need_lo || a->lo, "%s: split-fp16 view needs a lo plane", name);
    FT_REQUIRE(a->c % 8 == 0 && a->ld % 8 == 0 && a->coff % 8 == 0, "%s: c/ld/coff must be multiples of 8 (c=%d ld=%d coff=%d)", name, a->c, a->ld, a->coff);
    FT_REQUIRE(a->coff + a->c <= a->ld, "%s: view exceeds pixel stride", name);
    FT_REQUIRE(((uintptr_t)a->hi & 15) == 0 && ((uintptr_t)a->lo & 15) == 0, "%s: planes must be 16-byte aligned", name);
    return 0;
}

}  // namespace ft

using namespace ft;

extern "C" {

int ft_abi_version(void) { return FT_ABI_VERSION; }
int ft_sizeof(int which) { return which == 0 ? (int)sizeof(ft_act) : (int)sizeof(ft_conv_desc); }
const char* ft_last_error(void) { return ft::g_err; }
long long ft_launch_count(void) { return ft::g_launches.load(); }
void ft_reset_launch_count(void) { ft::g_launches = 0; }
int ft_set_debug(int which, int value) {
    FT_REQUIRE(which >= 0 && which < ft::FT_DBG_COUNT, "set_debug: unknown knob %d", which);
    ft::g_knobs[which].store(value);
    return 0;
}

int ft_correlation_shape(int H, int W, int pad, int ksize, int max_disp, int stride1, int stride2, int* outC, int* outH, int* outW) {
    FT_REQUIRE(ksize >= 1 && (ksize & 1) && stride1 >= 1 && stride2 >= 1 && max_disp >= 0 && pad >= 0, "correlation: bad parameters");
    const int border = (ksize - 1) / 2 + max_disp;
    const int d = max_disp / stride2;
    *outC = (2 * d + 1) * (2 * d + 1);
    *outH = (int)ceilf((float)(H + 2 * pad - 2 * border) / (float)stride1);
    *outW = (int)ceilf((float)(W + 2 * pad - 2 * border) / (float)stride1);
    FT_REQUIRE(*outH > 0 && *outW > 0, "correlation: empty output (%d x %d)", *outH, *outW);
    return 0;
}

int ft_correlation_forward(const float* in1, const float* in2, float* out, int B, int C, int H, int W, int pad, int ksize,
                           int max_disp, int stride1, int stride2, int fuse_leaky, void* stream) {
    FT_REQUIRE(in1 && in2 && out, "correlation: null pointer");
    FT_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "correlation: bad shape");
    int oc, oh, ow;
    if (int e = ft_correlation_shape(H, W, pad, ksize, max_disp, stride1, stride2, &oc, &oh, &ow)) return e;
    cudaStream_t s = (cudaStream_t)stream;
    if (ksize == 1 && stride1 == 1 && stride2 == 2 && max_disp == 20 && pad == 20) {
        dim3 grid(ceil_div(W, 64), ceil_div(H, CORR_R) * (21 / CORR_TJ), B);
        corr_fast_kernel<<<grid, CORR_THREADS, 0, s>>>(in1, in2, out, C, H, W, fuse_leaky);
        return check_launch("corr_fast_kernel");
    }
    dim3 block(32, 8), grid(ceil_div(ow, 32), oh, B);
    corr_generic_kernel<<<grid, block, 0, s>>>(in1, in2, out, C, H, W, oh, ow, pad, (ksize - 1) / 2, max_disp, stride1, stride2,
                                               max_disp / stride2, fuse_leaky);
    return check_launch("corr_generic_kernel");
}

int ft_resample2d_forward(const float* img, const float* flow, float* out, int B, int C, int H, int W, void* stream) {
    FT_REQUIRE(img && flow && out && B > 0 && C > 0 && H > 0 && W > 0, "resample2d: bad arguments");
    resample2d_kernel<<<dim3(ceil_div(W, PIX_THREADS), min(H, 65535), min(B, 65535)), PIX_THREADS, 0, (cudaStream_t)stream>>>(img, flow, out, B, C, H, W);
    return check_launch("resample2d_kernel");
}

int ft_channelnorm_forward(const float* x, float* out, int B, int C, int H, int W, void* stream) {
    FT_REQUIRE(x && out && B > 0 && C > 0 && H > 0 && W > 0, "channelnorm: bad arguments");
    channelnorm_kernel<<<dim3(ceil_div(W, PIX_THREADS), min(H, 65535), min(B, 65535)), PIX_THREADS, 0, (cudaStream_t)stream>>>(x, out, B, C, H, W);
    return check_launch("channelnorm_kernel");
}

int ft_maxpool3x3s2(const ft_act* in, const ft_act* out, void* stream) {
    if (int e = check_act(in, "maxpool.in", false)) return e;
    if (int e = check_act(out, "maxpool.out", false)) return e;
    FT_REQUIRE(out->h == (in->h + 1) / 2 && out->w == (in->w + 1) / 2 && out->n == in->n && out->c == in->c, "maxpool: shape mismatch");
    maxpool3x3s2_kernel<<<dim3(ceil_div(out->w * (out->c / 8), 256), min(out->h, 65535), min(out->n, 65535)), 256, 0, (cudaStream_t)stream>>>(*in, *out);
    return check_launch("maxpool3x3s2_kernel");
}

int ft_pack_image(const float* src, long long sb, long long sc, long long sh, long long sw, int C, float mul, const ft_act* dst, int xpad_l, int xpad_r,
                  void* stream) {
    if (int e = check_act(dst, "pack_image.dst", false)) return e;
    FT_REQUIRE(src && C > 0 && C <= dst->c, "pack_image: bad channel count %d (view has %d)", C, dst ? dst->c : -1);
    FT_REQUIRE(xpad_l >= 0 && xpad_r >= 0 && dst->w > xpad_l + xpad_r, "pack_image: bad x-padding %d/%d for a %d-pixel row", xpad_l, xpad_r, dst->w);
    const int W = dst->w - xpad_l - xpad_r;
    pack_image_kernel<<<grid_for((size_t)dst->n * dst->h * W * (dst->c / 8), 256), 256, 0, (cudaStream_t)stream>>>(src, sb, sc, sh, sw, C, mul, *dst, W, xpad_l);
    return check_launch("pack_image_kernel");
}

int ft_pack_image_s2d(const float* src, int B, int H, int W, const ft_act* dst, int xpad_l, int xpad_r, void* stream) {
    if (int e = check_act(dst, "pack_image_s2d.dst", false)) return e;
    FT_REQUIRE(src && B > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "pack_image_s2d: H and W must be even (got %d x %d)", H, W);
    FT_REQUIRE(((uintptr_t)src & 7) == 0, "pack_image_s2d: src must be 8-byte aligned");
    FT_REQUIRE(dst->c == 16 && dst->n == B && dst->h == H / 2 && xpad_l >= 0 && xpad_r >= 0 && dst->w == W / 2 + xpad_l + xpad_r,
               "pack_image_s2d: destination must be [B, H/2, xpad_l + W/2 + xpad_r, 16] (got n=%d h=%d w=%d c=%d)", dst->n, dst->h, dst->w, dst->c);
    pack_image_s2d_kernel<<<dim3(ceil_div(W / 2, PIX_THREADS), min(H / 2, 65535), min(B, 65535)), PIX_THREADS, 0, (cudaStream_t)stream>>>(src, B, H, W, *dst, xpad_l);
    return check_launch("pack_image_s2d_kernel");
}

int ft_pack_pair_ws_floats(int B) { return B * 3 * (2 * PAIR_MEAN_CHUNKS + 1); }

int ft_pack_pair(const float* pair, int B, int H, int W, float rgb_max, float* mean_ws, float* x_f32, const ft_act* dst, int xpad_l, int xpad_r,
                 void* stream) {
    FT_REQUIRE(pair && mean_ws && B > 0 && H > 0 && W > 0, "pack_pair: bad arguments");
    ft_act d = {};
    if (dst) {
        if (int e = check_act(dst, "pack_pair.dst", false)) return e;
        FT_REQUIRE(xpad_l >= 0 && xpad_r >= 0, "pack_pair: bad x-padding");
        FT_REQUIRE(dst->c >= 8 && dst->n == B && dst->h == H && dst->w == W + xpad_l + xpad_r, "pack_pair: dst shape mismatch");
        d = *dst;
    }
    FT_REQUIRE(((uintptr_t)mean_ws & 7) == 0, "pack_pair: workspace must be 8-byte aligned");
    double* partial = reinterpret_cast<double*>(mean_ws);
    float* mean = mean_ws + (size_t)B * 3 * PAIR_MEAN_CHUNKS * 2;
    pair_mean_partial_kernel<<<dim3(B * 3, PAIR_MEAN_CHUNKS), 256, 0, (cudaStream_t)stream>>>(pair, 2 * H * W, partial);
    if (check_launch("pair_mean_partial_kernel")) return 1;
    pair_mean_final_kernel<<<B * 3, 32, 0, (cudaStream_t)stream>>>(partial, 2 * H * W, mean);
    if (check_launch("pair_mean_final_kernel")) return 1;
    pack_pair_kernel<<<dim3(ceil_div(W, PIX_THREADS), min(H, 65535), min(B, 65535)), PIX_THREADS, 0, (cudaStream_t)stream>>>(pair, mean, B, H, W, rgb_max, x_f32, d, dst != nullptr, xpad_l);
    return check_launch("pack_pair_kernel");
}

int ft_pack_pair_u8(const uint8_t* frames, int B, int H, int W, int bgr, float rgb_max, void* sum_ws, float* x_f32, const ft_act* dst, int xpad_l,
                    int xpad_r, void* stream) {
    FT_REQUIRE(frames && sum_ws && B > 0 && H > 0 && W > 0, "pack_pair_u8: bad arguments");
    FT_REQUIRE(((size_t)H * W) % 4 == 0 && ((uintptr_t)frames & 3) == 0, "pack_pair_u8: H*W must be a multiple of 4 and the frames 4-byte aligned");
    FT_REQUIRE(B <= 65535 && ((uintptr_t)sum_ws & 7) == 0, "pack_pair_u8: at most 65535 pairs; the workspace must be 8-byte aligned");
    ft_act d = {};
    if (dst) {
        if (int e = check_act(dst, "pack_pair_u8.dst", false)) return e;
        FT_REQUIRE(xpad_l >= 0 && xpad_r >= 0, "pack_pair_u8: bad x-padding");
        FT_REQUIRE(dst->c >= 8 && dst->n == B && dst->h == H && dst->w == W + xpad_l + xpad_r, "pack_pair_u8: dst shape mismatch");
        d = *dst;
    }
    cudaStream_t s = (cudaStream_t)stream;
    if (cudaMemsetAsync(sum_ws, 0, sizeof(unsigned long long) * 3 * B, s) != cudaSuccess) {
        set_error("pack_pair_u8: memset of the workspace failed");
        return 1;
    }
    const size_t bytes_per_pair = (size_t)2 * H * W * 3;
    const int bx = min(ceil_div((int)(bytes_per_pair / 12), 256), max(1, device_sms() * 8 / B));
    pair_u8_sum_kernel<<<dim3(bx, B), 256, 0, s>>>(frames, bytes_per_pair, (unsigned long long*)sum_ws);
    if (check_launch("pair_u8_sum_kernel")) return 1;
    pack_pair_u8_kernel<<<dim3(ceil_div(W, 256), min(H, 65535), B), 256, 0, s>>>(frames, (const unsigned long long*)sum_ws, B, H, W, bgr != 0, rgb_max, x_f32, d,
                                                                                  dst != nullptr, xpad_l);
    return check_launch("pack_pair_u8_kernel");
}

int ft_upsample4x(const float* src, float* dst, int B, int C, int h, int w, float mul, int mode, void* stream) {
    FT_REQUIRE(src && dst && B > 0 && C > 0 && h > 0 && w > 0 && (mode == 0 || mode == 1), "upsample4x: bad arguments");
    const dim3 grid(ceil_div(w, 128), min(4 * h, 65535), min(B * C, 65535));
    const int vec = ((uintptr_t)dst & 15) == 0;
    if (mode == 1) upsample4x_kernel<1><<<grid, 128, 0, (cudaStream_t)stream>>>(src, dst, B * C, h, w, mul, vec);
    else upsample4x_kernel<0><<<grid, 128, 0, (cudaStream_t)stream>>>(src, dst, B * C, h, w, mul, vec);
    return check_launch("upsample4x_kernel");
}

int ft_warp_concat(const float* x6, const float* flow, float div_flow, int B, int H, int W, const ft_act* dst, void* stream) {
    if (int e = check_act(dst, "warp_concat.dst", false)) return e;
    FT_REQUIRE(x6 && flow && dst->c == 16 && dst->n == B && dst->h == H && dst->w == W, "warp_concat: bad arguments");
    warp_concat_kernel<<<dim3(ceil_div(W, PIX_THREADS), min(H, 65535), min(B, 65535)), PIX_THREADS, 0, (cudaStream_t)stream>>>(x6, flow, div_flow, B, H, W, *dst);
    return check_launch("warp_concat_kernel");
}

int ft_fusion_concat(const float* x6, const float* flow_sd, const float* flow_s2, int B, int H, int W, const ft_act* dst, void* stream) {
    if (int e = check_act(dst, "fusion_concat.dst", false)) return e;
    FT_REQUIRE(x6 && flow_sd && flow_s2 && dst->c == 16 && dst->n == B && dst->h == H && dst->w == W, "fusion_concat: bad arguments");
    fusion_concat_kernel<<<dim3(ceil_div(W, PIX_THREADS), min(H, 65535), min(B, 65535)), PIX_THREADS, 0, (cudaStream_t)stream>>>(x6, flow_sd, flow_s2, B, H, W, *dst);
    return check_launch("fusion_concat_kernel");
}

int ft_flow_head(const float* part, const float* bias, float* flow, int B, int H, int W, void* stream) {
    FT_REQUIRE(part && flow && B > 0 && H > 0 && W > 0, "flow_head: bad arguments");
    flow_head_kernel<<<dim3(ceil_div(W, 128), min(H, 65535), min(B, 65535)), 128, 0, (cudaStream_t)stream>>>(part, bias, flow, B, H, W);
    return check_launch("flow_head_kernel");
}

int ft_flow_upsample(const float* flow, const float* w, const float* bias, int B, int h, int wd, const ft_act* dst, void* stream) {
    if (int e = check_act(dst, "flow_upsample.dst", false)) return e;
    FT_REQUIRE(flow && w && B > 0 && h > 0 && wd > 0, "flow_upsample: bad arguments");
    FT_REQUIRE(dst->c == 8 && dst->n == B && dst->h == 2 * h && dst->w == 2 * wd, "flow_upsample: dst must be an 8-channel view of a [%d,%d,%d] map", B, 2 * h, 2 * wd);
    flow_upsample_kernel<<<dim3(ceil_div(2 * wd, 128), min(2 * h, 65535), min(B, 65535)), 128, 0, (cudaStream_t)stream>>>(flow, w, bias, B, h, wd, *dst);
    return check_launch("flow_upsample_kernel");
}

int ft_decode_keypoints(const float* hm, int N, int C, int h, int w, const float* center, const float* scale, int adjust, float threshold,
                        const int32_t* valid, float* kps, void* stream) {
    FT_REQUIRE(hm && center && scale && kps && N > 0 && C > 0 && h > 0 && w > 0, "decode_keypoints: bad arguments");
    decode_kernel<<<N * C, 256, 0, (cudaStream_t)stream>>>(hm, C, h, w, center, scale, adjust, threshold, valid, kps);
    return check_launch("decode_kernel");
}

int ft_box_propagation(const float* kps, const int32_t* frame_of, int N, int K, const float* flow, int B, int H, int W, float* boxes, void* stream) {
    FT_REQUIRE(kps && flow && boxes && N > 0 && K > 0 && B > 0 && H > 0 && W > 0, "box_propagation: bad arguments");
    boxprop_kernel<<<ceil_div(N, 4), 128, 0, (cudaStream_t)stream>>>(kps, frame_of, N, K, flow, B, H, W, boxes);
    return check_launch("boxprop_kernel");
}

}  // extern "C"
