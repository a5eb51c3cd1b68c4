/*
 * flowtrack_b200.h -- C ABI of libflowtrack_b200.so, the B200 (sm_100a) implementation of the
 * FlowTrack per-frame-pair inference hot path (PoseResNet forward, FlowNet2-family forward,
 * Correlation / Resample2d / ChannelNorm, heat-map decode and box propagation).
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name says `host`;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - nothing allocates, frees or synchronises: the calls only enqueue kernels on `stream`, so they are
 *     CUDA-graph-capture safe;  inputs are borrowed and never modified, outputs are owned by the caller;
 *   - return 0 on success, non-zero on an argument or launch error; ft_last_error() describes it
 *     (thread-local).  NB the reference glue returns 1 for success (correlation_cuda.c:90); the Python
 *     shim maps non-zero to RuntimeError so module-level behaviour (raise on failure) is unchanged.
 *
 * Reference interfaces replaced (paths relative to the reference repo root):
 *   ft_correlation_forward   Correlation_forward_cuda            lib/flownet/networks/correlation_package/src/correlation_cuda.h:1-9
 *                            (+ Correlation_forward_cuda_kernel   .../src/correlation_cuda_kernel.h:5-39; kernels .cu:10-106)
 *   ft_resample2d_forward    Resample2d_cuda_forward             lib/flownet/networks/resample2d_package/src/Resample2d_cuda.h:1
 *                            (kernel_Resample2d_updateOutput      .../src/Resample2d_kernel.cu:20-66,188-206)
 *   ft_channelnorm_forward   ChannelNorm_cuda_forward            lib/flownet/networks/channelnorm_package/src/ChannelNorm_cuda.h:1
 *                            (kernel_ChannelNorm_updateOutput     .../src/ChannelNorm_kernel.cu:19-51,82-98)
 *   ft_conv2d_fprop          every nn.Conv2d / nn.ConvTranspose2d(k4,s2,p1) + eval BatchNorm2d / bias +
 *                            ReLU / LeakyReLU(0.1) + residual add on the path (cuDNN/ATen calls made by
 *                            lib/pose/models/{resnet.py:19-27,blocks.py:89-120,pose_deconv.py:19-46} and
 *                            lib/flownet/networks/{submodules.py:7-38,FlowNetS.py,FlowNetC.py,FlowNetSD.py,FlowNetFusion.py})
 *   ft_maxpool3x3s2          nn.MaxPool2d(3,2,1)                  lib/pose/models/resnet.py:23
 *   ft_pack_image / ft_pack_pair   `.cuda()` + per-sample rgb-mean normalisation   lib/flownet/model/models.py:112-119,255-257
 *   ft_upsample4x            nn.Upsample(scale_factor=4, bilinear|nearest)          lib/flownet/networks/FlowNetS.py:58, models.py:32,59-60
 *   ft_decode_keypoints      max_preds + final_preds + transform_preds              lib/pose/utils/evaluation.py:11-35, transforms.py:173-228
 *   ft_box_propagation       box_propagation                                       lib/tracking/flow_utils.py:7-35
 *   ft_crop_affine           transform_image (cv2.warpAffine) per person box         lib/tracking/net_utils.py:48-57, lib/pose/utils/transforms.py:231-240
 *   ft_flip_w / ft_flip_merge  flip test: np.flip + swaplr_image + average                  tools/pose/main.py:289-299, lib/pose/utils/transforms.py:140-170
 *   ft_nms_oks               nms_oks + compute_oks                                   lib/pose/utils/evaluation.py:60-101
 *   ft_flow_to_image         flowlib.flow_to_image (Middlebury colour coding)         lib/flownet/utils/flowlib.py:243-282,430-521 (tools/flownet/demo.py:93)
 *   ft_pack_pair_u8          cv2 uint8 frames -> RGB float pair + normalisation      lib/tracking/net_utils.py:83-87, lib/flownet/model/models.py:112-119
 *   ft_track_merge / ft_track_select   detect()'s merge + pose_est's box -> centre/scale   lib/tracking/net_utils.py:25-33,49-50, tools/tracking/demo.py:35-42
 *   ft_nms                   pth_nms / cpu_nms / nms_kernel                          lib/detection/nms/pth_nms.py:5-46, src/nms.c:36-62, src/cuda/nms_kernel.cu:13-70
 */
#ifndef FLOWTRACK_B200_H_
#define FLOWTRACK_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FT_ABI_VERSION 18

/* ---- activations ------------------------------------------------------------------------------
 * Inter-layer activations live in HBM as "split-fp16" NHWC planes: value = float(hi) + float(lo),
 * hi = fp16(v), lo = fp16(v - hi)  (22 significant bits).  `hi` and `lo` are
 * two [n,h,w,ld] __half arrays; a view selects channels [coff, coff+c) of every pixel, so producers
 * write straight into slices of a channel-concatenated buffer (torch.cat on the path costs nothing).
 * c, ld and coff are multiples of 8 (16-byte vectors).  lo == NULL means "hi only" (fast mode). */
typedef struct {
    void* hi;
    void* lo;
    int32_t n, h, w;
    int32_t c;      /* channels visible through the view (padded to a multiple of 8)            */
    int32_t ld;     /* elements between consecutive pixels                                       */
    int32_t coff;   /* first channel of the view inside a pixel                                  */
} ft_act;

enum { FT_ACT_NONE = 0, FT_ACT_RELU = 1, FT_ACT_LEAKY = 2 };
enum { FT_CONV = 0, FT_DECONV_K4S2P1 = 1 };
enum { FT_PREC_SPLIT = 0,  /* 3 tcgen05 MMAs per tile (hi*hi + lo*hi + hi*lo): fp32-grade results  */
       FT_PREC_FAST  = 1   /* 1 MMA on the hi planes only: fp16-operand results                    */ };
enum { FT_K_PACKED = 0,    /* k = tap*in.c + ci, k_pad = round_up(taps*in.c, 64): im2col gathered by cp.async;
                              for the small-Cin stems (a 64-channel TMA box per tap would be mostly padding)   */
       FT_K_TAP64  = 1,    /* k = tap*cin_k + ci with cin_k = round_up(in.c, 64): one TMA box per (tap, 64 ch)  */
       FT_K_ROW    = 2     /* k = r*64 + s*in.c + ci (kernel row r, column s; kw*in.c <= 64, k_pad = kh*64): one TMA
                              box per kernel row -- the kw taps of a row are contiguous in an NHWC image, so a
                              tensor map whose pixel stride is smaller than its 128-byte box row reads them in one
                              go (the 7x7 stems).  Needs in_xpad_l/r zero columns in the input rows.             */ };
enum { FT_IMPL_TCGEN05 = 0,  /* tcgen05/TMEM implicit GEMM (the product path)                      */
       FT_IMPL_SIMT    = 1   /* fp32 CUDA-core kernel, same data formats (bring-up / cross-check)  */ };

/* One fused convolution:  out = act( conv(in, W) * scale[c] + shift[c] (+ residual) ).
 * Weights are supplied pre-transformed by the host (flowtrack/pytorch_b200/engine.py):
 *   w_gemm : fp32 [n_phase][k_pad][cout_pad], k = tap * in.c + ci  (taps row-major over the kernel
 *            window; for FT_DECONV_K4S2P1 the 4 output-parity phases each see a 2x2 tap subset);
 *   w_tile : the same matrix as tcgen05 operand tiles: for each (phase, n_tile, k_block of 64)
 *            two [block_n][64] fp16 K-major 128B-swizzled images (hi then lo), contiguous; after the
 *            k_pad/64 main blocks of a (phase, n_tile) come `res_blocks` residual blocks (see below).
 * scale/shift have cout_pad entries (zeros in the padding).                                        */
typedef struct {
    ft_act in;
    ft_act out;             /* out.hi == NULL -> no split-fp16 output                              */
    ft_act res;             /* res.hi == NULL -> no residual                                       */
    float* out_f32;         /* optional contiguous NCHW fp32 output [n, cout, ho, wo]              */
    const void*  w_tile;
    const float* w_gemm;
    const float* scale;
    const float* shift;
    int32_t kind;           /* FT_CONV | FT_DECONV_K4S2P1                                          */
    int32_t kh, kw, stride, pad;
    int32_t cout;           /* valid output channels                                               */
    int32_t cout_pad;       /* multiple of block_n                                                 */
    int32_t k_pad;          /* multiple of 64                                                      */
    int32_t block_n;        /* 16, 32, 64, 128 or 256                                              */
    int32_t act;            /* FT_ACT_*                                                            */
    float   slope;          /* LeakyReLU slope                                                     */
    int32_t precision;      /* FT_PREC_*                                                           */
    int32_t impl;           /* FT_IMPL_*                                                           */
    int32_t k_mode;         /* FT_K_*: K layout of w_gemm / w_tile                                 */
    int32_t in_xpad_l;      /* the rows of `in` physically start with in_xpad_l and end with in_xpad_r */
    int32_t in_xpad_r;      /* zero pixels: in.w counts them, the image is in.w - l - r wide (0/0 normally) */
    int32_t cta_pair;       /* 1: run 256-pixel x block_n tiles on CTA pairs (tcgen05 cta_group::2, clusters of 2);
                               needs FT_K_TAP64 / FT_K_ROW and block_n >= 64.
                               2: clusters of four = two CTA pairs on the same 256 pixels and two neighbouring n-tiles;
                               the activation (A) tile is fetched once per cluster -- each CTA loads half of it and
                               TMA-multicasts it to its twin in the other pair -- which cuts the L2 -> SM traffic of a
                               k-block from 48 to 32 KiB per SM.  FT_K_TAP64, block_n 64 / 128, even n-tile count.   */
    int32_t res_blocks;     /* residual k-blocks appended to every (phase, n_tile) of w_tile:
                               ceil(block_n/64) when the layer is called with a residual, else 0.
                               Block j holds diag(f) for channels [64j, 64j+64) of the tile, where
                               f[c] = 1/scale[c] must be a power of two (the host folds BatchNorm's
                               scale into the weights of such layers), so acc += res * f exactly.   */
    int32_t pad_hi;         /* zero padding after the last row / column (<= pad); -1 = same as `pad`.  The space-to-depth form
                               of a 7x7 stride-2 stem is a 4x4 stride-1 convolution padded 2 before and 1 after.       */
    int32_t tile_order;     /* 0: the persistent CTAs walk the output tiles first image -> last image, 1: last -> first.
                               A layer that runs in the opposite order of its producer starts on the part of its
                               input that is still in the 126 MB L2 (engine.py alternates the order along a plan). */
    int32_t epilogue_buffers; /* 0 / 1: one staging slab per epilogue warp group; 2: two (ping-pong: a group fills one slab while
                               the TMA store of its previous slab still reads the other).  2 costs 32 KiB of shared memory =
                               one pipeline stage, and pays on the small-K 1x1 layers whose time is the epilogue's store
                               hand-off chain.  block_n 64 / 128, TMA variants only; engine.Plan.autotune() chooses.        */
    int32_t splitk;         /* <= 1: off.  n: the K range of every tile is cut into n parts that run as separate work items on
                               different CTAs (a layer with fewer tiles than SMs -- layer3/4 of a small batch, every layer of a
                               strong-scaling shard -- then still fills the GPU).  Parts park their fp32 partial sums in
                               splitk_ws; the CTA whose part arrives last adds them in part order (deterministic) and runs the
                               normal epilogue.  FT_K_TAP64 / FT_K_ROW only.                                                  */
    int32_t splitk_counters_len;   /* entries of splitk_counters: >= number of CTA tiles (tiles x 2 for cta_pair)             */
    void*   splitk_ws;      /* 16-byte aligned scratch, >= CTA tiles x parts x block_n x 128 x 4 bytes                        */
    uint64_t splitk_ws_bytes;
    void*   splitk_counters;/* uint32 per CTA tile, zero before the first launch (the kernel leaves them zero)               */
    int32_t ctas_per_sm;    /* 0 / 1: one persistent CTA per SM.  2: two (half the shared memory each: 2 pipeline stages, 16-channel
                               staging slabs), so one CTA's epilogue and pipeline fill run under the other's main loop -- for the
                               layers bound by the per-tile epilogue latency chain (stems, 64-channel 3x3 convolutions and
                               deconvolutions).  block_n 64 on CTA pairs only; engine.Plan.autotune() chooses.               */
    int32_t reserved0;
} ft_conv_desc;

int         ft_abi_version(void);
/* sizeof(ft_act) (which = 0) / sizeof(ft_conv_desc) (which = 1): lets a binding verify its struct layout. */
int         ft_sizeof(int which);
const char* ft_last_error(void);
/* Number of kernels this library has launched since load / since the last reset (bench accounting). */
long long   ft_launch_count(void);
void        ft_reset_launch_count(void);
/* Development knobs for the profiling experiments under tools/dev (0 = skip mask of the conv / correlation kernels,
 * 1 = accumulation chunk, 2 = correlation parts, 3 = correlation kernel variant, 4 = no programmatic dependent launch,
 * 5 = cap on the persistent grid of the conv kernel).
 * All zero by default; the library
 * never reads the environment.  Results under a non-zero skip mask are garbage by design.                          */
int         ft_set_debug(int which, int value);

/* ---- the three native ops of the reference (contiguous NCHW fp32, as there) --------------------- */
int ft_correlation_shape(int H, int W, int pad, int ksize, int max_disp, int stride1, int stride2,
                         int* outC, int* outH, int* outW);
/* in1,in2: [B,C,H,W] -> out: [B,outC,outH,outW].  fuse_leaky != 0 applies LeakyReLU(0.1) (FlowNetC.py:87). */
int ft_correlation_forward(const float* in1, const float* in2, float* out, int B, int C, int H, int W,
                           int pad, int ksize, int max_disp, int stride1, int stride2, int fuse_leaky,
                           void* stream);
/* Tensor-core cost volume (ft_corr.cu) for the FlowNetC geometry (pad 20, kernel 1, max displacement 20, stride1 1,
 * stride2 2; FlowNetC.py:33): the four stride-2 parity classes are independent stride-1 correlations, computed as
 * banded GEMMs by tcgen05 MMAs.  ft_correlation_tc_supported() != 0 when the geometry and C (64..256, multiple of 64)
 * fit that kernel; maps wider than 64 pixels (frames wider than 512) are tiled in 32-column strips of every parity class.
 *   ft_correlation_act        in1/in2: split-fp16 NHWC activations (the conv stack's layout) -> out: fp32 NCHW
 *                             [B,441,H,W]; precision = FT_PREC_SPLIT (fp32-grade) | FT_PREC_FAST (fp16 operands).
 *   ft_correlation_forward_ws the reference's signature (fp32 NCHW in, like ft_correlation_forward) plus a
 *                             caller-owned scratch buffer -- the counterpart of the rInput1/rInput2 scratch tensors
 *                             of Correlation_forward_cuda (correlation_cuda.c:36-42): the inputs are re-laid out
 *                             into it (NHWC split-fp16) and ft_correlation_act runs on the copies.  Geometries the
 *                             tensor-core kernel does not cover fall through to ft_correlation_forward.
 *                             workspace: 256-byte aligned, >= ft_correlation_workspace_bytes(B,C,H,W).          */
int    ft_correlation_tc_supported(int C, int H, int W, int pad, int ksize, int max_disp, int stride1, int stride2);
int    ft_correlation_act(const ft_act* in1, const ft_act* in2, float* out, int fuse_leaky, int precision, void* stream);
/* Same kernel, pipeline-internal output layout: parity-planar fp32 [B,441,4,H/2,W/2] -- class py*2+px holds the pixels
 * (2a+py, 2b+px) as a dense [H/2,W/2] plane, so a warp store is 128 contiguous bytes (in NCHW the stride-2 stores fill
 * half of every 32-byte sector).  ft_pack_corr_planar converts it into channels [0,441) of a split-fp16 NHWC view
 * (FlowNetC's 473-channel merge buffer, FlowNetC.py:101); H and W must be even.                                      */
int    ft_correlation_act_planar(const ft_act* in1, const ft_act* in2, float* out, int fuse_leaky, int precision, void* stream);
int    ft_pack_corr_planar(const float* planar, const ft_act* dst, void* stream);
size_t ft_correlation_workspace_bytes(int B, int C, int H, int W);
int    ft_correlation_forward_ws(const float* in1, const float* in2, float* out, int B, int C, int H, int W,
                                 int pad, int ksize, int max_disp, int stride1, int stride2, int fuse_leaky,
                                 int precision, void* workspace, size_t workspace_bytes, void* stream);
/* img: [B,C,H,W], flow: [B,2,H,W] (u,v) -> out: [B,C,H,W] */
int ft_resample2d_forward(const float* img, const float* flow, float* out, int B, int C, int H, int W, void* stream);
/* x: [B,C,H,W] -> out: [B,1,H,W] */
int ft_channelnorm_forward(const float* x, float* out, int B, int C, int H, int W, void* stream);

/* ---- conv stack ---------------------------------------------------------------------------------- */
int ft_conv2d_fprop(const ft_conv_desc* d, void* stream);
int ft_maxpool3x3s2(const ft_act* in, const ft_act* out, void* stream);
/* src: [B,C,H,W] fp32 (arbitrary strides given in elements) -> split-fp16 NHWC view `dst` channels
 * [0,C), multiplied by `mul`; channels [C, dst->c) are zero-filled.  dst->w = xpad_l + W + xpad_r: the
 * image is written at column xpad_l of every destination row, the padding columns are left untouched
 * (the caller zero-fills them once; FT_K_ROW convolutions read them as the convolution's zero border). */
int ft_pack_image(const float* src, long long sb, long long sc, long long sh, long long sw, int C,
                  float mul, const ft_act* dst, int xpad_l, int xpad_r, void* stream);
/* Space-to-depth packing for 3-channel 7x7 stride-2 stems: src [B,3,H,W] fp32 contiguous (H, W even) -> split-fp16 NHWC
 * view dst [B, H/2, xpad_l + W/2 + xpad_r, 16]: channel (py*2+px)*4 + c holds src[b, c, 2Y+py, 2X+px], the 4th channel of
 * every sub-pixel is zero.  The stem then runs as a 4x4 stride-1 convolution (pad 2, pad_hi 1) whose kernel row is exactly
 * one 64-element k-block: K = 256 instead of 448, and half the packed bytes.                                            */
int ft_pack_image_s2d(const float* src, int B, int H, int W, const ft_act* dst, int xpad_l, int xpad_r, void* stream);
/* FlowNet2 input normalisation: pair [B,3,2,H,W] (RGB 0..255) -> x = (pair - mean_bc) / rgb_max laid
 * out as [B,6,H,W] fp32 (x_f32, optional) and as split-fp16 NHWC (dst, optional, 6 -> 8 channels).
 * `mean_ws` is an 8-byte aligned workspace of ft_pack_pair_ws_floats(B) floats (partial sums + means). */
int ft_pack_pair_ws_floats(int B);
int ft_pack_pair(const float* pair, int B, int H, int W, float rgb_max, float* mean_ws, float* x_f32,
                 const ft_act* dst, int xpad_l, int xpad_r, void* stream);   /* dst->w = xpad_l + W + xpad_r, as above */
/* The same from uint8 frames as the reference's callers hold them (cv2 frames, lib/tracking/net_utils.py:83-87): frames
 * [B,2,H,W,3] uint8, memory channel order BGR (bgr != 0, cv2) or RGB; the BGR->RGB flip, the float conversion, the per-sample
 * colour mean over both frames (exact integer sums) and the division by rgb_max happen in the packing pass, so a frame pair
 * crosses PCIe as 6 bytes per pixel instead of 24.  sum_ws: B*3 x 8 bytes of scratch (cleared by a memset node).          */
int ft_pack_pair_u8(const uint8_t* frames, int B, int H, int W, int bgr, float rgb_max, void* sum_ws, float* x_f32,
                    const ft_act* dst, int xpad_l, int xpad_r, void* stream);
/* src: [B,C,h,w] fp32 -> dst: [B,C,4h,4w] fp32 times `mul`.  mode 0 = bilinear (align_corners=False),
 * 1 = nearest.                                                                                        */
int ft_upsample4x(const float* src, float* dst, int B, int C, int h, int w, float mul, int mode, void* stream);
/* FlowNet2 stacking glue (models.py:126-131): warped = resample2d(x[:,3:6], flow); err = |x[:,0:3]-warped|_2;
 * writes [x(6) | warped(3) | flow/div_flow(2) | err(1)] as 12 (+4 zero) split-fp16 NHWC channels.     */
int ft_warp_concat(const float* x6, const float* flow, float div_flow, int B, int H, int W,
                   const ft_act* dst, void* stream);
/* FlowNet2 fusion input (models.py:150-172): [x0(3) | sd(2) | s2(2) | |sd| | |s2| | err_sd | err_s2]
 * as 11 (+5 zero) split-fp16 NHWC channels; sd / s2 are the two full-resolution flow candidates.      */
int ft_fusion_concat(const float* x6, const float* flow_sd, const float* flow_s2, int B, int H, int W,
                     const ft_act* dst, void* stream);

/* Flow prediction head (lib/flownet/networks/submodules.py:20-21, predict_flow = Conv2d(C, 2, 3, 1, 1)).  The path
 * runs the 3x3 two-channel convolution as a 1x1 convolution with 18 outputs through ft_conv2d_fprop (out_f32),
 * part[n, (r*3+s)*2 + c, y, x] = W[c,:,r,s] . in[n,:,y,x], and this call finishes it:
 * flow[n,c,y,x] = bias[c] + sum_{r,s} part[n, (r*3+s)*2 + c, y+r-1, x+s-1].  part: [B,18,H,W], flow: [B,2,H,W] fp32. */
int ft_flow_head(const float* part, const float* bias, float* flow, int B, int H, int W, void* stream);
/* upsampled_flowX_to_Y = ConvTranspose2d(2, 2, 4, 2, 1) (FlowNetS.py:41-44): flow [B,2,h,w] fp32, w = the module's
 * weight [2][2][4][4], bias [2] or NULL -> split-fp16 NHWC channels [0,2) (+6 zeros) of the 8-channel view dst
 * at [B,2h,2w].                                                                                          */
int ft_flow_upsample(const float* flow, const float* w, const float* bias, int B, int h, int wd,
                     const ft_act* dst, void* stream);

/* ---- decode -------------------------------------------------------------------------------------- */
/* hm: [N,C,h,w] fp32; center: [N,2], scale: [N] fp32 -> kps: [N,C,3] = (x, y, score) in image
 * coordinates (max_preds + optional quarter-pixel shift + inverse crop transform).  `threshold` is
 * max_preds' argument (evaluation.py:11,18-19): heat-map coordinates are zeroed unless score > threshold
 * (final_preds always passes 0).  valid: optional [N] int32; persons with valid == 0 (empty slots of the
 * tracking pipeline) get all-zero records.                                                           */
int ft_decode_keypoints(const float* hm, int N, int C, int h, int w, const float* center,
                        const float* scale, int adjust, float threshold, const int32_t* valid, float* kps,
                        void* stream);
/* kps: [N,K,3] previous-frame key-points; frame_of: [N] int32 index of each person's frame pair;
 * flow: [B,2,H,W] -> boxes: [N,4] (x0,y0,x1,y1).  Key-point pixels are clamped into the flow field
 * before the look-up (the reference indexes unclamped, lib/tracking/flow_utils.py:23-24).             */
int ft_box_propagation(const float* kps, const int32_t* frame_of, int N, int K, const float* flow,
                       int B, int H, int W, float* boxes, void* stream);

/* ---- tracking glue either side of the networks ------------------------------ */
/* Person crops of pose_est (lib/tracking/net_utils.py:48-57 -> transforms.py:231-240 transform_image ->
 * cv2.warpAffine, bilinear, zero border), every box of a batch in one launch.  frames: N_f images addressed by
 * element strides (sn frame, sc channel, sh row, sw column), uint8 (is_u8 != 0: the reference's cv2 frames,
 * bit-exact with cv2.warpAffine's fixed-point bilinear path) or float32; centers [N,2] / scales [N] as in pose_est
 * (box centre, max(box_h, box_w/out_w*out_h)); frame_of [N] int32 (NULL: frame 0) -> out [N,C,out_h,out_w] fp32.
 * reverse_channels != 0 writes channel c to C-1-c (RGB frames of the flow input -> the BGR crops pose_est feeds). */
int ft_crop_affine(const void* frames, int is_u8, long long sn, long long sc, long long sh, long long sw,
                   int H, int W, int C, const float* centers, const float* scales, const int32_t* frame_of,
                   int N, int out_h, int out_w, int reverse_channels, float* out, void* stream);
/* The same crops (3-channel frames, even crop size) written straight into the space-to-depth split-fp16 layout of the PoseResNet
 * stem (see ft_pack_image_s2d): dst [N, out_h/2, xpad_l + out_w/2 + xpad_r, 16].  The fp32 crops never touch HBM.          */
int ft_crop_affine_s2d(const void* frames, int is_u8, long long sn, long long sc, long long sh, long long sw,
                       int H, int W, int C, const float* centers, const float* scales, const int32_t* frame_of,
                       int N, int out_h, int out_w, int reverse_channels, const ft_act* dst, int xpad_l, int xpad_r,
                       void* stream);
/* Flip test (tools/pose/main.py:289-299; swaplr_image, lib/pose/utils/transforms.py:140-170).
 * ft_flip_w mirrors the last axis: src/dst [rows, W] fp32 (a [B,C,H,W] batch is rows = B*C*H).
 * ft_flip_merge: out[n,c,y,x] = 0.5 * (hm[n,c,y,x] + hm_flipped[n, perm[c], y, W-1-x]); perm [C] int32 is the left/right
 * joint permutation (get_pairs, transforms.py:60-96); `out` may alias `hm`.                                       */
int ft_flip_w(const float* src, float* dst, long long rows, int W, void* stream);
int ft_flip_merge(const float* hm, const float* hm_flipped, const int32_t* perm, int N, int C, int H, int W,
                  float* out, void* stream);
/* Greedy box NMS per frame (net_utils.py:29-32 -> detection/nms/pth_nms.py:5-46).  dets [frames,N,5] =
 * (x1,y1,x2,y2,score), counts [frames] valid boxes per frame (NULL: N), N <= 1024.  inclusive != 0 suppresses at
 * IoU >= thresh (cpu_nms, src/nms.c:57 -- the variant the tracking path reaches with its CPU tensor), 0 at
 * IoU > thresh (nms_kernel.cu:62).  keep [frames,N] int32: indices of the kept boxes in descending score order
 * (-1 beyond num_keep[frame]).                                                                        */
int ft_nms(const float* dets, const int32_t* counts, int frames, int N, float thresh, int inclusive,
           int32_t* keep, int32_t* num_keep, void* stream);

/* The tracker's device-resident glue between box propagation and pose_est (tools/tracking/demo.py:35-42 process_frame;
 * lib/tracking/net_utils.py:25-33 detect, :49-50 pose_est).  F frames (streams), P person slots per frame, D detector boxes.
 *   ft_track_merge   cand [F,D+P,5] = per frame [det (det_count[f], NULL = D) | prop_boxes with the scores prev_dets[:,4]
 *                    (prev_count[f])], compact, cand_count [F]: np.concatenate((dets, prop_dets)) of net_utils.py:27-29.
 *   (ft_nms over cand / cand_count)
 *   ft_track_select  the first min(num_keep[f], P) survivors, highest score first -> dets [F,P,5], count [F], and for every slot
 *                    the crop centre [F*P,2] / scale [F*P] of pose_est (box centre; max(box_h, box_w/out_w*out_h)) and
 *                    valid [F*P] int32; empty slots get a whole-frame crop and valid = 0.                                 */
int ft_track_merge(const float* prop_boxes, const float* prev_dets, const int32_t* prev_count, const float* det,
                   const int32_t* det_count, int F, int P, int D, float* cand, int32_t* cand_count, void* stream);
int ft_track_select(const float* cand, const int32_t* keep, const int32_t* num_keep, int F, int N, int P, int out_h, int out_w,
                    int frame_h, int frame_w, float* dets, int32_t* count, float* centers, float* scales, int32_t* valid,
                    void* stream);

/* OKS-NMS over poses (nms_oks + compute_oks(ground_truth=False), lib/pose/utils/evaluation.py:60-101).  joints
 * [frames,N,K,3] = (x, y, score), scores [frames,N] (pose score), areas [frames,N], delta [K] (per-joint falloff),
 * counts [frames] or NULL, N <= 1024.  A pose suppresses a lower-scored one when their OKS (mean over the joints both
 * score >= kpt_thresh, scale = mean of the two areas) exceeds oks_thresh.  keep / num_keep as in ft_nms.            */
int ft_nms_oks(const float* joints, const float* scores, const float* areas, const float* delta, const int32_t* counts,
               int frames, int N, int K, float oks_thresh, float kpt_thresh, int32_t* keep, int32_t* num_keep,
               void* stream);

/* Middlebury colour coding of flow fields (flowlib.flow_to_image, lib/flownet/utils/flowlib.py:243-282 with compute_color
 * :430-471 and make_color_wheel :474-521; tools/flownet/demo.py:93).  flow [B,2,H,W] fp32 (the networks' output layout)
 * -> img [B,H,W,3] uint8 RGB, each flow normalised by its own largest radius; entries with |u| or |v| > 1e7 are "unknown"
 * and painted black.  maxrad_ws: B x 4 bytes of scratch (cleared by a memset node, then two kernels).                 */
int ft_flow_to_image(const float* flow, int B, int H, int W, void* maxrad_ws, uint8_t* img, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FLOWTRACK_B200_H_ */
