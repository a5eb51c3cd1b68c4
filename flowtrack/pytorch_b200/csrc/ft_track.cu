// ft_track.cu -- the tracking glue either side of the two networks, sm_100a only:
//   ft_crop_affine  person crops for pose_est (lib/tracking/net_utils.py:48-57 -> lib/pose/utils/transforms.py:231-240
//                   transform_image -> cv2.warpAffine(INTER_LINEAR, BORDER_CONSTANT 0)), all boxes of a batch in one launch
//   ft_flip_w / ft_flip_merge  the flip test of tools/pose/main.py:289-299 (mirror the crops, run them in the same batch, mirror the
//                   heat-maps back with the left/right joints swapped and average)
//   ft_nms          greedy person-box NMS (lib/tracking/net_utils.py:29-32 -> lib/detection/nms/pth_nms.py:5-46,
//                   src/nms.c:36-62 / src/cuda/nms_kernel.cu:13-70), one CTA per frame, nothing leaves the GPU
#include "ft_common.cuh"

namespace ft {

// ------------------------------------------------------------------------------------------------
// Affine crops.  cv2.warpAffine is the reference's third-party arithmetic (OpenCV, un-vendored); its bilinear path
// works on a fixed-point grid, restated here (modules/imgproc/src/imgwarp.cpp, WarpAffineInvoker):
//   M = 2x3 forward map, inverted in double precision (dst -> src);
//   X = (round(M0*x*1024) + round((M1*y + M2)*1024) + 16) >> 5     (1/32-pixel source coordinate), Y likewise;
//   sx = X >> 5, fx = X & 31;  the four taps outside the image count as 0;
//   8-bit images: out = (sum_i w_i * p_i + 512) >> 10 with w = (32-fx)(32-fy), fx(32-fy), (32-fx)fy, fx*fy  (exact
//   integer form of OpenCV's 15-bit weight table); float images: the same weights / 1024 in floating point.
// The forward map is get_transform(center, scale, res) (transforms.py:173-184, factor 1, rot 0):
//   M0 = M4 = res_h / scale,  M2 = -res_h*cx/scale + res_w/2,  M5 = -res_h*cy/scale + res_h/2.
// ------------------------------------------------------------------------------------------------
struct CropSrc {
    const void* base;
    long long sn, sc, sh, sw;   // element strides: frame, channel, row, column
    int H, W, C, is_u8;
};

// inverse of get_transform(center, scale, (oh, ow)) as cv::warpAffine inverts it: src = (M0*x + M1*y + M2, M3*x + M4*y + M5)
struct CropMap {
    double M0, M1, M2, M3, M4, M5;
};
__device__ __forceinline__ CropMap crop_map(double cx, double cy, double scale, int oh, int ow) {
    CropMap m;
    m.M0 = (double)oh / scale, m.M1 = 0.0, m.M3 = 0.0, m.M4 = (double)oh / scale;
    m.M2 = -(double)oh * cx / scale + 0.5 * ow, m.M5 = -(double)oh * cy / scale + 0.5 * oh;
    double D = m.M0 * m.M4 - m.M1 * m.M3;
    D = D != 0 ? 1. / D : 0;
    const double A11 = m.M4 * D, A22 = m.M0 * D;
    m.M0 = A11, m.M1 *= -D, m.M3 *= -D, m.M4 = A22;
    const double b1 = -m.M0 * m.M2 - m.M1 * m.M5, b2 = -m.M3 * m.M2 - m.M4 * m.M5;
    m.M2 = b1, m.M5 = b2;
    return m;
}
// the C channel values of output pixel (x, y) of a crop (fixed-point bilinear taps as described above)
template <int MAXC>
__device__ __forceinline__ void crop_pixel(const CropSrc& src, const CropMap& m, long long f, int x, int y, float (&v)[MAXC]) {
    const int X = (__double2int_rn(m.M0 * x * 1024.0) + __double2int_rn((m.M1 * y + m.M2) * 1024.0) + 16) >> 5;
    const int Y = (__double2int_rn(m.M3 * x * 1024.0) + __double2int_rn((m.M4 * y + m.M5) * 1024.0) + 16) >> 5;
    const int sx = max(-32768, min(32767, X >> 5)), sy = max(-32768, min(32767, Y >> 5));
    const int fx = X & 31, fy = Y & 31;
    const int w00 = (32 - fx) * (32 - fy), w01 = fx * (32 - fy), w10 = (32 - fx) * fy, w11 = fx * fy;
    const bool x0 = sx >= 0 && sx < src.W, x1 = sx + 1 >= 0 && sx + 1 < src.W;
    const bool y0 = sy >= 0 && sy < src.H, y1 = sy + 1 >= 0 && sy + 1 < src.H;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        if (c >= src.C) break;
        const long long b = f * src.sn + (long long)c * src.sc;
        const long long o00 = b + sy * src.sh + sx * src.sw, o01 = o00 + src.sw, o10 = o00 + src.sh, o11 = o10 + src.sw;
        if (src.is_u8) {
            const uint8_t* p = (const uint8_t*)src.base;
            const int s = w00 * (y0 && x0 ? (int)p[o00] : 0) + w01 * (y0 && x1 ? (int)p[o01] : 0) + w10 * (y1 && x0 ? (int)p[o10] : 0) +
                          w11 * (y1 && x1 ? (int)p[o11] : 0);
            v[c] = (float)((s + 512) >> 10);
        } else {
            const float* p = (const float*)src.base;
            const float k = 1.f / 1024.f;
            v[c] = (y0 && x0 ? __ldg(p + o00) : 0.f) * (w00 * k) + (y0 && x1 ? __ldg(p + o01) : 0.f) * (w01 * k) + (y1 && x0 ? __ldg(p + o10) : 0.f) * (w10 * k) +
                   (y1 && x1 ? __ldg(p + o11) : 0.f) * (w11 * k);
        }
    }
}

__global__ void __launch_bounds__(256) crop_affine_kernel(CropSrc src, const float* __restrict__ centers, const float* __restrict__ scales,
                                                          const int32_t* __restrict__ frame_of, int N, int oh, int ow, int reverse_c,
                                                          float* __restrict__ out) {
    const int n = blockIdx.z;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    __shared__ CropMap sm;                      // one double-precision inversion per block instead of one per thread
    if (threadIdx.x == 0) sm = crop_map(centers[2 * n], centers[2 * n + 1], scales[n], oh, ow);
    __syncthreads();
    if (x >= ow || y >= oh) return;
    const CropMap m = sm;
    float v[4];
    crop_pixel<4>(src, m, frame_of ? frame_of[n] : 0, x, y, v);
    for (int c = 0; c < src.C; ++c) {
        const int co = reverse_c ? src.C - 1 - c : c;
        out[(((size_t)n * src.C + co) * oh + y) * ow + x] = v[c];
    }
}

// The same crops written straight into the space-to-depth split-fp16 layout the PoseResNet stem reads (ft_pack_image_s2d):
// dst [N, oh/2, xl + ow/2 + xr, 16], channel (py*2+px)*4 + c = crop[c, 2Y+py, 2X+px], 4th channel of every sub-pixel zero.
// One thread per s2d pixel: 4 crop pixels x 3 channels; the fp32 crops never touch HBM (38 MB written + read per 64 crops).
__global__ void __launch_bounds__(256) crop_affine_s2d_kernel(CropSrc src, const float* __restrict__ centers, const float* __restrict__ scales,
                                                              const int32_t* __restrict__ frame_of, int N, int oh, int ow, int reverse_c, ft_act d, int xl) {
    const int n = blockIdx.z;
    const int X = blockIdx.x * 32 + (threadIdx.x & 31), Y = blockIdx.y * 8 + (threadIdx.x >> 5);
    __shared__ CropMap sm;                      // one double-precision inversion per block instead of one per thread
    if (threadIdx.x == 0) sm = crop_map(centers[2 * n], centers[2 * n + 1], scales[n], oh, ow);
    __syncthreads();
    if (X >= (ow >> 1) || Y >= (oh >> 1)) return;
    const CropMap m = sm;
    const long long f = frame_of ? frame_of[n] : 0;
    float o[16];
#pragma unroll
    for (int sp = 0; sp < 4; ++sp) {
        float v[3];
        crop_pixel<3>(src, m, f, 2 * X + (sp & 1), 2 * Y + (sp >> 1), v);
        o[sp * 4 + 0] = reverse_c ? v[2] : v[0], o[sp * 4 + 1] = v[1], o[sp * 4 + 2] = reverse_c ? v[0] : v[2], o[sp * 4 + 3] = 0.f;
    }
    const size_t pix = ((size_t)n * d.h + Y) * d.w + X + xl;
    __half* hi = reinterpret_cast<__half*>(d.hi) + pix * d.ld + d.coff;
    __half* lo = d.lo ? reinterpret_cast<__half*>(d.lo) + pix * d.ld + d.coff : nullptr;
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        Half8 h8, l8;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            __half h0, l0, h1, l1;
            split_f32(o[g * 8 + 2 * k], h0, l0);
            split_f32(o[g * 8 + 2 * k + 1], h1, l1);
            h8.v[k] = __halves2half2(h0, h1), l8.v[k] = __halves2half2(l0, l1);
        }
        *reinterpret_cast<Half8*>(hi + g * 8) = h8;
        if (lo) *reinterpret_cast<Half8*>(lo + g * 8) = l8;
    }
}

// ------------------------------------------------------------------------------------------------
// NMS.  One CTA per frame: bitonic sort of (score, index) in shared memory, the upper-triangular suppression matrix as
// bit rows in shared memory (row i: which lower-scored boxes box i suppresses), then one warp walks the sorted list
// (lane w owns word w of the 1024-bit "removed" set).  IoU uses the reference's +1 pixel convention.
// ------------------------------------------------------------------------------------------------
constexpr int NMS_MAX = 1024;
constexpr int NMS_WORDS = NMS_MAX / 32;

__global__ void __launch_bounds__(1024) nms_kernel(const float* __restrict__ dets, const int32_t* __restrict__ counts, int N, float thresh, int inclusive,
                                                   int32_t* __restrict__ keep, int32_t* __restrict__ num_keep) {
    extern __shared__ uint32_t nms_smem[];
    float* key = reinterpret_cast<float*>(nms_smem);                  // [1024] scores (sorted descending)
    int* idx = reinterpret_cast<int*>(nms_smem + NMS_MAX);            // [1024] original indices
    float* box = reinterpret_cast<float*>(nms_smem + 2 * NMS_MAX);    // [1024][4] sorted boxes
    uint32_t* mask = nms_smem + 6 * NMS_MAX;                          // [n][words]
    const int f = blockIdx.x, t = threadIdx.x;
    const int n = counts ? min(max(counts[f], 0), N) : N;
    const float* d = dets + (size_t)f * N * 5;
    int32_t* kp = keep + (size_t)f * N;
    int np2 = 32;
    while (np2 < n) np2 <<= 1;
    // sort by score, descending; ties by ascending index; padding sorts last
    key[t] = t < n ? d[t * 5 + 4] : -INFINITY;
    idx[t] = t < n ? t : 0x7fffffff;
    __syncthreads();
    for (int k = 2; k <= np2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            const int p = t ^ j;
            if (t < np2 && p > t) {
                const bool desc = (t & k) == 0;
                const float a = key[t], b = key[p];
                const int ia = idx[t], ib = idx[p];
                const bool a_first = a > b || (a == b && ia < ib);      // a belongs before b in the final order
                if (desc ? !a_first : a_first) key[t] = b, key[p] = a, idx[t] = ib, idx[p] = ia;
            }
            __syncthreads();
        }
    if (t < n) {
        const float* s = d + (size_t)idx[t] * 5;
        box[4 * t] = s[0], box[4 * t + 1] = s[1], box[4 * t + 2] = s[2], box[4 * t + 3] = s[3];
    }
    __syncthreads();
    const int words = (n + 31) >> 5;
    // suppression bits: thread t handles row t, all words (n*words IoU bit-words in total)
    if (t < n) {
        const float ax1 = box[4 * t], ay1 = box[4 * t + 1], ax2 = box[4 * t + 2], ay2 = box[4 * t + 3];
        const float sa = (ax2 - ax1 + 1) * (ay2 - ay1 + 1);
        for (int w = 0; w < words; ++w) {
            uint32_t bits = 0;
            if (w >= (t >> 5)) {
                for (int b = 0; b < 32; ++b) {
                    const int j = w * 32 + b;
                    if (j <= t || j >= n) continue;
                    const float bx1 = box[4 * j], by1 = box[4 * j + 1], bx2 = box[4 * j + 2], by2 = box[4 * j + 3];
                    const float iw = fmaxf(fminf(ax2, bx2) - fmaxf(ax1, bx1) + 1, 0.f), ih = fmaxf(fminf(ay2, by2) - fmaxf(ay1, by1) + 1, 0.f);
                    const float inter = iw * ih;
                    const float sb = (bx2 - bx1 + 1) * (by2 - by1 + 1);
                    const float ovr = inter / (sa + sb - inter);
                    if (inclusive ? ovr >= thresh : ovr > thresh) bits |= 1u << b;
                }
            }
            mask[t * words + w] = bits;
        }
    }
    __syncthreads();
    if (t < 32) {
        uint32_t removed = 0;          // lane w: word w of the removed set
        int kept = 0;
        for (int i = 0; i < n; ++i) {
            const uint32_t wv = __shfl_sync(0xffffffffu, removed, i >> 5);
            if (!((wv >> (i & 31)) & 1u)) {
                if (t == 0) kp[kept] = idx[i];
                ++kept;
                if (t < words) removed |= mask[i * words + t];
            }
        }
        if (t == 0) num_keep[f] = kept;
        for (int i = kept + t; i < N; i += 32) kp[i] = -1;
    }
}

// OKS-NMS over poses (nms_oks / compute_oks(ground_truth=False), lib/pose/utils/evaluation.py:60-101): same one-CTA-per-frame
// scheme as the box NMS; pose i suppresses a lower-scored pose j when
//   mean_{k: score_i[k] >= t and score_j[k] >= t} exp(-|p_i[k]-p_j[k]|^2 / (2 delta[k]^2 ((area_i+area_j)/2 + eps))) > oks_thresh
// (no counted joint: OKS = 0).  Arithmetic in double like the numpy reference.
__global__ void __launch_bounds__(1024) nms_oks_kernel(const float* __restrict__ joints, const float* __restrict__ scores, const float* __restrict__ areas,
                                                       const float* __restrict__ delta, const int32_t* __restrict__ counts, int N, int K, float oks_thresh,
                                                       float kpt_thresh, int32_t* __restrict__ keep, int32_t* __restrict__ num_keep) {
    extern __shared__ uint32_t nms_smem[];
    float* key = reinterpret_cast<float*>(nms_smem);
    int* idx = reinterpret_cast<int*>(nms_smem + NMS_MAX);
    uint32_t* mask = nms_smem + 2 * NMS_MAX;
    const int f = blockIdx.x, t = threadIdx.x;
    const int n = counts ? min(max(counts[f], 0), N) : N;
    const float* J = joints + (size_t)f * N * K * 3;
    const float* A = areas + (size_t)f * N;
    int32_t* kp = keep + (size_t)f * N;
    int np2 = 32;
    while (np2 < n) np2 <<= 1;
    key[t] = t < n ? scores[(size_t)f * N + t] : -INFINITY;
    idx[t] = t < n ? t : 0x7fffffff;
    __syncthreads();
    for (int k = 2; k <= np2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            const int p = t ^ j;
            if (t < np2 && p > t) {
                const bool desc = (t & k) == 0;
                const float a = key[t], b = key[p];
                const int ia = idx[t], ib = idx[p];
                const bool a_first = a > b || (a == b && ia < ib);
                if (desc ? !a_first : a_first) key[t] = b, key[p] = a, idx[t] = ib, idx[p] = ia;
            }
            __syncthreads();
        }
    const int words = (n + 31) >> 5;
    if (t < n) {
        const float* ji = J + (size_t)idx[t] * K * 3;
        const double ai = A[idx[t]];
        for (int w = 0; w < words; ++w) {
            uint32_t bits = 0;
            if (w >= (t >> 5)) {
                for (int b = 0; b < 32; ++b) {
                    const int j = w * 32 + b;
                    if (j <= t || j >= n) continue;
                    const float* jj = J + (size_t)idx[j] * K * 3;
                    const double scale = (ai + (double)A[idx[j]]) / 2 + 2.220446049250313e-16;
                    double sum = 0;
                    int cnt = 0;
                    for (int k = 0; k < K; ++k) {
                        if (ji[3 * k + 2] >= kpt_thresh && jj[3 * k + 2] >= kpt_thresh) {
                            const double dx = (double)ji[3 * k] - (double)jj[3 * k], dy = (double)ji[3 * k + 1] - (double)jj[3 * k + 1];
                            const double d2 = (double)delta[k] * (double)delta[k];
                            sum += exp(-(dx * dx + dy * dy) / 2 / d2 / scale);
                            ++cnt;
                        }
                    }
                    const double oks = cnt ? sum / cnt : 0.0;
                    if (oks > (double)oks_thresh) bits |= 1u << b;
                }
            }
            mask[t * words + w] = bits;
        }
    }
    __syncthreads();
    if (t < 32) {
        uint32_t removed = 0;
        int kept = 0;
        for (int i = 0; i < n; ++i) {
            const uint32_t wv = __shfl_sync(0xffffffffu, removed, i >> 5);
            if (!((wv >> (i & 31)) & 1u)) {
                if (t == 0) kp[kept] = idx[i];
                ++kept;
                if (t < words) removed |= mask[i * words + t];
            }
        }
        if (t == 0) num_keep[f] = kept;
        for (int i = kept + t; i < N; i += 32) kp[i] = -1;
    }
}

// ------------------------------------------------------------------------------------------------
// Flip test (tools/pose/main.py:289-299; swaplr_image / swap, lib/pose/utils/transforms.py:140-170).
// ------------------------------------------------------------------------------------------------
__global__ void flip_w_kernel(const float* __restrict__ src, float* __restrict__ dst, long long rows, int W) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < rows * W; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / W;
        const int x = (int)(i - r * W);
        dst[i] = __ldg(src + r * W + (W - 1 - x));
    }
}

// out[n,c,y,x] = 0.5 * (hm[n,c,y,x] + hm_flipped[n, perm[c], y, W-1-x])
__global__ void flip_merge_kernel(const float* __restrict__ hm, const float* __restrict__ hm_flipped, const int32_t* __restrict__ perm, int N, int C, int H, int W,
                                  float* __restrict__ out) {
    const long long total = (long long)N * C * H * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W);
        const long long r = i / W;
        const int y = (int)(r % H);
        const long long nc = r / H;
        const int c = (int)(nc % C);
        const long long n = nc / C;
        const float f = __ldg(hm_flipped + ((n * C + perm[c]) * H + y) * W + (W - 1 - x));
        out[i] = (__ldg(hm + i) + f) * 0.5f;
    }
}


// ------------------------------------------------------------------------------------------------
// Middlebury flow colour coding (flow_to_image / compute_color / make_color_wheel, lib/flownet/utils/flowlib.py:243-282,
// 430-521) -- the visualisation step after the flow forward (tools/flownet/demo.py:92-99).  Arithmetic follows numpy's
// promotion on the reference's expressions for a float32 flow: radius and its per-image maximum in float32 (separately
// rounded multiply / add / sqrt), everything after the normalisation in double.
// ------------------------------------------------------------------------------------------------
constexpr float FLOWVIS_UNKNOWN = 1e7f;

__global__ void flowvis_max_kernel(const float* __restrict__ flow, int HW, uint32_t* __restrict__ maxrad_bits) {
    const int b = blockIdx.y;
    const float* u = flow + (size_t)b * 2 * HW;
    const float* v = u + HW;
    float m = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
        float uu = u[i], vv = v[i];
        if (fabsf(uu) > FLOWVIS_UNKNOWN || fabsf(vv) > FLOWVIS_UNKNOWN) uu = 0.f, vv = 0.f;
        const float r = __fsqrt_rn(__fadd_rn(__fmul_rn(uu, uu), __fmul_rn(vv, vv)));
        if (r > m) m = r;                                      // NaN compares false, as in np.max over a NaN-free field
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(maxrad_bits + b, __float_as_uint(m));   // non-negative floats order like their bit patterns
}

// make_color_wheel (flowlib.py:474-521): RY 15, YG 6, GC 4, CB 11, BM 13, MR 6 = 55 colours, ramps floor(255*i/n)
__constant__ uint8_t c_flow_wheel[55][3] = {{255,0,0}, {255,17,0}, {255,34,0}, {255,51,0}, {255,68,0}, {255,85,0}, {255,102,0}, {255,119,0}, {255,136,0}, {255,153,0}, {255,170,0}, {255,187,0}, {255,204,0}, {255,221,0}, {255,238,0}, {255,255,0}, {213,255,0}, {170,255,0}, {128,255,0}, {85,255,0}, {43,255,0}, {0,255,0}, {0,255,63}, {0,255,127}, {0,255,191}, {0,255,255}, {0,232,255}, {0,209,255}, {0,186,255}, {0,163,255}, {0,140,255}, {0,116,255}, {0,93,255}, {0,70,255}, {0,47,255}, {0,24,255}, {0,0,255}, {19,0,255}, {39,0,255}, {58,0,255}, {78,0,255}, {98,0,255}, {117,0,255}, {137,0,255}, {156,0,255}, {176,0,255}, {196,0,255}, {215,0,255}, {235,0,255}, {255,0,255}, {255,0,213}, {255,0,170}, {255,0,128}, {255,0,85}, {255,0,43}};

__global__ void flowvis_color_kernel(const float* __restrict__ flow, int HW, const uint32_t* __restrict__ maxrad_bits, uint8_t* __restrict__ img) {
    const int b = blockIdx.y;
    const float* u = flow + (size_t)b * 2 * HW;
    const float* v = u + HW;
    const double denom = (double)__uint_as_float(maxrad_bits[b]) + 2.220446049250313e-16;      // maxrad + np.finfo(float).eps
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
        float uu = u[i], vv = v[i];
        uint8_t* o = img + ((size_t)b * HW + i) * 3;
        const bool unknown = fabsf(uu) > FLOWVIS_UNKNOWN || fabsf(vv) > FLOWVIS_UNKNOWN;
        if (unknown || isnan(uu) || isnan(vv)) {
            o[0] = o[1] = o[2] = 0;
            continue;
        }
        const double un = (double)uu / denom, vn = (double)vv / denom;
        const double rad = sqrt(un * un + vn * vn);
        const double a = atan2(-vn, -un) / 3.141592653589793;
        const double fk = (a + 1.0) / 2.0 * 54.0 + 1.0;
        const int k0 = (int)floor(fk);
        int k1 = k0 + 1;
        if (k1 == 56) k1 = 1;
        const double f = fk - (double)k0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double c0 = (double)c_flow_wheel[k0 - 1][c] / 255.0, c1 = (double)c_flow_wheel[k1 - 1][c] / 255.0;
            double col = (1.0 - f) * c0 + f * c1;
            col = rad <= 1.0 ? 1.0 - rad * (1.0 - col) : col * 0.75;
            o[c] = (uint8_t)floor(255.0 * col);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// The tracker's glue between box propagation and pose_est (tools/tracking/demo.py:35-42 process_frame;
// lib/tracking/net_utils.py:25-33 detect, :49-50 pose_est), device-resident so the recurrent state (person boxes and
// key-points of the previous frame) never leaves the GPU:
//   merge : candidates of frame f = [detector boxes (det_count[f]) | propagated boxes with the previous scores (prev_count[f])]
//           (np.concatenate((dets, prop_dets)), net_utils.py:27-29; prop_dets = [prop_boxes | prev_dets[:,4]], demo.py:38)
//   (ft_nms over the candidates)
//   select: the first min(num_keep, P) survivors become the frame's person slots: dets [F,P,5], count [F], and per slot the crop
//           centre / scale of pose_est (centre = box centre, scale = max(box_h, box_w / out_w * out_h), computed in double as
//           numpy does) and a valid flag; empty slots get a harmless whole-frame crop and valid = 0.
// ------------------------------------------------------------------------------------------------
__global__ void track_merge_kernel(const float* __restrict__ prop_boxes, const float* __restrict__ prev_dets, const int32_t* __restrict__ prev_count,
                                   const float* __restrict__ det, const int32_t* __restrict__ det_count, int F, int P, int D,
                                   float* __restrict__ cand, int32_t* __restrict__ cand_count) {
    const int N = D + P;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F * N) return;
    const int f = i / N, j = i - f * N;
    const int nd = det ? min(max(det_count ? det_count[f] : D, 0), D) : 0;
    const int np = min(max(prev_count[f], 0), P);
    float* o = cand + (size_t)i * 5;
    if (j < nd) {
        const float* d = det + ((size_t)f * D + j) * 5;
        o[0] = d[0], o[1] = d[1], o[2] = d[2], o[3] = d[3], o[4] = d[4];
    } else if (j < nd + np) {
        const int k = j - nd;
        const float* b = prop_boxes + ((size_t)f * P + k) * 4;
        o[0] = b[0], o[1] = b[1], o[2] = b[2], o[3] = b[3], o[4] = prev_dets[((size_t)f * P + k) * 5 + 4];
    } else {
        o[0] = o[1] = o[2] = o[3] = 0.f, o[4] = -INFINITY;
    }
    if (j == 0) cand_count[f] = nd + np;
}

__global__ void track_select_kernel(const float* __restrict__ cand, const int32_t* __restrict__ keep, const int32_t* __restrict__ num_keep, int F, int N,
                                    int P, int out_h, int out_w, int frame_h, int frame_w, float* __restrict__ dets, int32_t* __restrict__ count,
                                    float* __restrict__ centers, float* __restrict__ scales, int32_t* __restrict__ valid) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F * P) return;
    const int f = i / P, k = i - f * P;
    const int n = min(max(num_keep[f], 0), P);
    float* o = dets + (size_t)i * 5;
    if (k < n) {
        const float* c = cand + ((size_t)f * N + keep[(size_t)f * N + k]) * 5;
        o[0] = c[0], o[1] = c[1], o[2] = c[2], o[3] = c[3], o[4] = c[4];
        centers[2 * i] = (float)(((double)c[0] + (double)c[2]) / 2.0);
        centers[2 * i + 1] = (float)(((double)c[1] + (double)c[3]) / 2.0);
        const double bh = (double)c[3] - (double)c[1], bw = ((double)c[2] - (double)c[0]) / (double)out_w * (double)out_h;
        scales[i] = (float)(bh > bw ? bh : bw);
        valid[i] = 1;
    } else {
        o[0] = o[1] = o[2] = o[3] = o[4] = 0.f;
        centers[2 * i] = 0.5f * frame_w, centers[2 * i + 1] = 0.5f * frame_h;
        scales[i] = (float)frame_h;
        valid[i] = 0;
    }
    if (k == 0) count[f] = n;
}

}  // namespace ft

using namespace ft;

extern "C" int ft_track_merge(const float* prop_boxes, const float* prev_dets, const int32_t* prev_count, const float* det, const int32_t* det_count,
                              int F, int P, int D, float* cand, int32_t* cand_count, void* stream) {
    FT_REQUIRE(F > 0 && P > 0 && D >= 0 && D + P <= NMS_MAX, "track_merge: bad shape (F=%d P=%d D=%d; at most %d candidates per frame)", F, P, D, NMS_MAX);
    FT_REQUIRE(prop_boxes && prev_dets && prev_count && cand && cand_count && (D == 0 || det), "track_merge: null pointer");
    track_merge_kernel<<<ceil_div(F * (D + P), 128), 128, 0, (cudaStream_t)stream>>>(prop_boxes, prev_dets, prev_count, D ? det : nullptr, det_count, F, P, D, cand,
                                                                                  cand_count);
    return check_launch("track_merge_kernel");
}

extern "C" int ft_track_select(const float* cand, const int32_t* keep, const int32_t* num_keep, int F, int N, int P, int out_h, int out_w, int frame_h,
                               int frame_w, float* dets, int32_t* count, float* centers, float* scales, int32_t* valid, void* stream) {
    FT_REQUIRE(F > 0 && P > 0 && N >= P && out_h > 0 && out_w > 0 && frame_h > 0 && frame_w > 0, "track_select: bad shape");
    FT_REQUIRE(cand && keep && num_keep && dets && count && centers && scales && valid, "track_select: null pointer");
    track_select_kernel<<<ceil_div(F * P, 128), 128, 0, (cudaStream_t)stream>>>(cand, keep, num_keep, F, N, P, out_h, out_w, frame_h, frame_w, dets, count, centers,
                                                                             scales, valid);
    return check_launch("track_select_kernel");
}

extern "C" int ft_flow_to_image(const float* flow, int B, int H, int W, void* maxrad_ws, uint8_t* img, void* stream) {
    FT_REQUIRE(B >= 0 && H > 0 && W > 0, "flow_to_image: bad shape");
    if (B == 0) return 0;
    FT_REQUIRE(flow && maxrad_ws && img, "flow_to_image: null pointer");
    FT_REQUIRE(B <= 65535, "flow_to_image: at most 65535 flows per call");
    const int HW = H * W;
    if (cudaMemsetAsync(maxrad_ws, 0, sizeof(uint32_t) * B, (cudaStream_t)stream) != cudaSuccess) {
        set_error("flow_to_image: memset of the workspace failed");
        return 1;
    }
    const int bx = min(ceil_div(HW, 256), device_sms() * 4);
    flowvis_max_kernel<<<dim3(bx, B), 256, 0, (cudaStream_t)stream>>>(flow, HW, (uint32_t*)maxrad_ws);
    if (check_launch("flowvis_max_kernel")) return 1;
    flowvis_color_kernel<<<dim3(bx, B), 256, 0, (cudaStream_t)stream>>>(flow, HW, (const uint32_t*)maxrad_ws, img);
    return check_launch("flowvis_color_kernel");
}

extern "C" int ft_crop_affine(const void* frames, int is_u8, long long sn, long long sc, long long sh, long long sw, int H, int W, int C,
                              const float* centers, const float* scales, const int32_t* frame_of, int N, int out_h, int out_w, int reverse_channels,
                              float* out, void* stream) {
    FT_REQUIRE(H > 0 && W > 0 && C > 0 && C <= 4 && out_h > 0 && out_w > 0 && N >= 0, "crop_affine: bad shape");
    if (N == 0) return 0;                    // no boxes in this batch: nothing to launch (the pointers may be null)
    FT_REQUIRE(frames && centers && scales && out, "crop_affine: null pointer");
    FT_REQUIRE(N <= 65535, "crop_affine: at most 65535 crops per call");
    CropSrc s = {frames, sn, sc, sh, sw, H, W, C, is_u8 != 0};
    crop_affine_kernel<<<dim3(ceil_div(out_w, 32), ceil_div(out_h, 8), N), 256, 0, (cudaStream_t)stream>>>(s, centers, scales, frame_of, N, out_h, out_w,
                                                                                                            reverse_channels != 0, out);
    return check_launch("crop_affine_kernel");
}

extern "C" int ft_crop_affine_s2d(const void* frames, int is_u8, long long sn, long long sc, long long sh, long long sw, int H, int W, int C,
                                  const float* centers, const float* scales, const int32_t* frame_of, int N, int out_h, int out_w, int reverse_channels,
                                  const ft_act* dst, int xpad_l, int xpad_r, void* stream) {
    FT_REQUIRE(H > 0 && W > 0 && C == 3 && out_h > 0 && out_w > 0 && out_h % 2 == 0 && out_w % 2 == 0 && N >= 0, "crop_affine_s2d: bad shape (3 channels, even crop size)");
    if (N == 0) return 0;
    FT_REQUIRE(frames && centers && scales && dst && dst->hi, "crop_affine_s2d: null pointer");
    FT_REQUIRE(N <= 65535, "crop_affine_s2d: at most 65535 crops per call");
    FT_REQUIRE(dst->c == 16 && dst->ld % 8 == 0 && dst->coff % 8 == 0 && dst->coff + 16 <= dst->ld && dst->n == N && dst->h == out_h / 2 &&
                   dst->w == xpad_l + out_w / 2 + xpad_r && xpad_l >= 0 && xpad_r >= 0,
               "crop_affine_s2d: dst must be a 16-channel view of [%d, %d, %d + %d + %d]", N, out_h / 2, xpad_l, out_w / 2, xpad_r);
    CropSrc s = {frames, sn, sc, sh, sw, H, W, C, is_u8 != 0};
    crop_affine_s2d_kernel<<<dim3(ceil_div(out_w / 2, 32), ceil_div(out_h / 2, 8), N), 256, 0, (cudaStream_t)stream>>>(s, centers, scales, frame_of, N, out_h, out_w,
                                                                                                                    reverse_channels != 0, *dst, xpad_l);
    return check_launch("crop_affine_s2d_kernel");
}

extern "C" int ft_nms(const float* dets, const int32_t* counts, int frames, int N, float thresh, int inclusive, int32_t* keep, int32_t* num_keep,
                      void* stream) {
    FT_REQUIRE(dets && keep && num_keep && frames > 0, "nms: null pointer / no frames");
    FT_REQUIRE(N >= 0 && N <= NMS_MAX, "nms: at most %d boxes per frame (got %d)", NMS_MAX, N);
    const int words = (N + 31) / 32;
    const int smem = (6 * NMS_MAX + (N > 0 ? N * words : 1)) * 4;
    static bool attr_set[MAX_DEVICES];
    if (ensure_dyn_smem(nms_kernel, (6 * NMS_MAX + NMS_MAX * NMS_WORDS) * 4, attr_set, "nms_kernel")) return 1;
    nms_kernel<<<frames, 1024, smem, (cudaStream_t)stream>>>(dets, counts, N, thresh, inclusive != 0, keep, num_keep);
    return check_launch("nms_kernel");
}

extern "C" int ft_flip_w(const float* src, float* dst, long long rows, int W, void* stream) {
    FT_REQUIRE(rows >= 0 && W > 0, "flip_w: bad shape");
    if (rows == 0) return 0;
    FT_REQUIRE(src && dst && src != dst, "flip_w: null pointer / in-place call");
    const long long total = rows * W;
    const int grid = (int)((total + 255) / 256 < (long long)device_sms() * 16 ? (total + 255) / 256 : (long long)device_sms() * 16);
    flip_w_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, dst, rows, W);
    return check_launch("flip_w_kernel");
}

extern "C" int ft_flip_merge(const float* hm, const float* hm_flipped, const int32_t* perm, int N, int C, int H, int W, float* out, void* stream) {
    FT_REQUIRE(N >= 0 && C > 0 && H > 0 && W > 0, "flip_merge: bad shape");
    if (N == 0) return 0;
    FT_REQUIRE(hm && hm_flipped && perm && out, "flip_merge: null pointer");
    const long long total = (long long)N * C * H * W;
    const int grid = (int)((total + 255) / 256 < (long long)device_sms() * 16 ? (total + 255) / 256 : (long long)device_sms() * 16);
    flip_merge_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(hm, hm_flipped, perm, N, C, H, W, out);
    return check_launch("flip_merge_kernel");
}

extern "C" int ft_nms_oks(const float* joints, const float* scores, const float* areas, const float* delta, const int32_t* counts, int frames, int N, int K,
                          float oks_thresh, float kpt_thresh, int32_t* keep, int32_t* num_keep, void* stream) {
    FT_REQUIRE(joints && scores && areas && delta && keep && num_keep && frames > 0 && K > 0, "nms_oks: null pointer / bad shape");
    FT_REQUIRE(N >= 0 && N <= NMS_MAX, "nms_oks: at most %d poses per frame (got %d)", NMS_MAX, N);
    const int words = (N + 31) / 32;
    const int smem = (2 * NMS_MAX + (N > 0 ? N * words : 1)) * 4;
    static bool attr_set[MAX_DEVICES];
    if (ensure_dyn_smem(nms_oks_kernel, (2 * NMS_MAX + NMS_MAX * NMS_WORDS) * 4, attr_set, "nms_oks_kernel")) return 1;
    nms_oks_kernel<<<frames, 1024, smem, (cudaStream_t)stream>>>(joints, scores, areas, delta, counts, N, K, oks_thresh, kpt_thresh, keep, num_keep);
    return check_launch("nms_oks_kernel");
}
