/*
 * oracle/ops_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the three native FlowNet2 ops of the reference
 * (simochen/flowtrack.pytorch).  The reference ships CUDA-only kernels for
 * these (its CPU entry points are stubs, correlation_package/src/correlation.c:3-33),
 * so this file restates the arithmetic of the .cu files on the host:
 *
 *   oracle_correlation_forward  <- correlation_package/src/correlation_cuda_kernel.cu:10-32 (zero padding,
 *                                  channel-last scratch) and :34-106 (the product/sum loop), with the
 *                                  output-shape math of correlation_package/src/correlation_cuda.c:25-34.
 *   oracle_resample2d_forward   <- resample2d_package/src/Resample2d_kernel.cu:20-66
 *   oracle_channelnorm_forward  <- channelnorm_package/src/ChannelNorm_kernel.cu:19-51
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  All tensors are contiguous NCHW float32.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* Threads the library will use (1 unless built with -fopenmp). */
int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Output shape of the correlation layer (correlation_cuda.c:22-34). */
void oracle_correlation_shape(int H, int W, int pad, int ksize, int max_disp, int stride1, int stride2,
                              int* outC, int* outH, int* outW) {
    int kernel_radius = (ksize - 1) / 2;
    int border = kernel_radius + max_disp;
    int pH = H + 2 * pad, pW = W + 2 * pad;
    int d = max_disp / stride2;
    *outC = (2 * d + 1) * (2 * d + 1);
    *outH = (int)ceilf((float)(pH - 2 * border) / (float)stride1);
    *outW = (int)ceilf((float)(pW - 2 * border) / (float)stride1);
}

/*
 * Cost volume.  in1, in2: [B,C,H,W]; out: [B,outC,outH,outW].
 * Follows the kernel's index math: padded channel-last copies (K1), then for each
 * output pixel and displacement the k*k*C products are accumulated in the same
 * 32-lane strided order as the kernel (lane c sums channels c, c+32, ...; then the
 * 32 partial sums are added serially, correlation_cuda_kernel.cu:80-97) so the
 * float rounding matches the GPU reference as closely as a host can.
 */
int oracle_correlation_forward(const float* in1, const float* in2, float* out,
                               int B, int C, int H, int W,
                               int pad, int ksize, int max_disp, int stride1, int stride2) {
    int outC, outH, outW;
    oracle_correlation_shape(H, W, pad, ksize, max_disp, stride1, stride2, &outC, &outH, &outW);
    const int pH = H + 2 * pad, pW = W + 2 * pad;
    const int krad = (ksize - 1) / 2;
    const int drad = max_disp / stride2;
    const int dsize = 2 * drad + 1;
    const size_t pimg = (size_t)pH * pW * C;
    float* r1 = (float*)calloc((size_t)B * pimg, sizeof(float));
    float* r2 = (float*)calloc((size_t)B * pimg, sizeof(float));
    if (!r1 || !r2) { free(r1); free(r2); return 0; }
    /* K1: channels_first -> zero padded [B, pH, pW, C] */
    for (int n = 0; n < B; ++n)
        for (int c = 0; c < C; ++c)
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) {
                    size_t src = (((size_t)n * C + c) * H + y) * W + x;
                    size_t dst = (size_t)n * pimg + ((size_t)(y + pad) * pW + (x + pad)) * C + c;
                    r1[dst] = in1[src];
                    r2[dst] = in2[src];
                }
    const float nelems = (float)(ksize * ksize * C);
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < B; ++n)
        for (int by = 0; by < outH; ++by)
            for (int bx = 0; bx < outW; ++bx) {
                const int y1 = by * stride1 + max_disp + krad;
                const int x1 = bx * stride1 + max_disp + krad;
                for (int tj = -drad; tj <= drad; ++tj)
                    for (int ti = -drad; ti <= drad; ++ti) {
                        float lane[32];
                        for (int l = 0; l < 32; ++l) lane[l] = 0.f;
                        const int x2 = x1 + ti * stride2, y2 = y1 + tj * stride2;
                        for (int j = -krad; j <= krad; ++j)
                            for (int i = -krad; i <= krad; ++i) {
                                const float* p1 = r1 + (size_t)n * pimg + ((size_t)(y1 + j) * pW + (x1 + i)) * C;
                                const float* p2 = r2 + (size_t)n * pimg + ((size_t)(y2 + j) * pW + (x2 + i)) * C;
                                for (int ch = 0; ch < C; ++ch) lane[ch & 31] += p1[ch] * p2[ch];
                            }
                        float s = 0.f;
                        for (int l = 0; l < 32; ++l) s += lane[l];
                        const int tc = (tj + drad) * dsize + (ti + drad);
                        out[(((size_t)n * outC + tc) * outH + by) * outW + bx] = s / nelems;
                    }
            }
    free(r1);
    free(r2);
    return 1;
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/*
 * Bilinear backward warp.  img: [B,C,H,W] (sampled), flow: [B,2,H,W] (u=x, v=y), out: [B,C,H,W].
 * Corner indices are clamped to the image, the weights come from the un-clamped
 * fractional parts (Resample2d_kernel.cu:42-53).  The reference multiplies with
 * double literals (`1. - alpha`), so the four products are formed in double and
 * each is added into a float accumulator, as there.
 */
int oracle_resample2d_forward(const float* img, const float* flow, float* out,
                              int B, int C, int H, int W) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c)
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) {
                    float dx = flow[(((size_t)b * 2 + 0) * H + y) * W + x];
                    float dy = flow[(((size_t)b * 2 + 1) * H + y) * W + x];
                    float xf = (float)x + dx, yf = (float)y + dy;
                    float alpha = xf - floorf(xf), beta = yf - floorf(yf);
                    int xL = clampi((int)floorf(xf), 0, W - 1);
                    int xR = clampi((int)(floorf(xf) + 1), 0, W - 1);
                    int yT = clampi((int)floorf(yf), 0, H - 1);
                    int yB = clampi((int)(floorf(yf) + 1), 0, H - 1);
                    const float* p = img + ((size_t)b * C + c) * H * W;
                    float val = 0.0f;
                    val += (1. - alpha) * (1. - beta) * p[(size_t)yT * W + xL];
                    val += (alpha) * (1. - beta) * p[(size_t)yT * W + xR];
                    val += (1. - alpha) * (beta)*p[(size_t)yB * W + xL];
                    val += (alpha) * (beta)*p[(size_t)yB * W + xR];
                    out[(((size_t)b * C + c) * H + y) * W + x] = val;
                }
    return 1;
}

/* L2 norm over channels.  x: [B,C,H,W] -> out: [B,1,H,W]  (ChannelNorm_kernel.cu:41-50). */
int oracle_channelnorm_forward(const float* x, float* out, int B, int C, int H, int W) {
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b)
        for (int y = 0; y < H; ++y)
            for (int xx = 0; xx < W; ++xx) {
                float r = 0.0f;
                for (int c = 0; c < C; ++c) {
                    float v = x[(((size_t)b * C + c) * H + y) * W + xx];
                    r += v * v;
                }
                out[((size_t)b * H + y) * W + xx] = sqrtf(r);
            }
    return 1;
}
