// ft_common.cuh -- shared helpers for libflowtrack_b200 (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../../include/flowtrack_b200.h"

namespace ft {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

// Returns 0 if the last launch was accepted, else records the CUDA error and returns 1.
int check_launch(const char* what);

// ---- per-device state ------------------------------------------------------------------------------
// One process may drive several GPUs (cudaSetDevice between calls): function attributes are per device, and so is
// the SM count the persistent grids are sized with.
constexpr int MAX_DEVICES = 64;
int current_device();          // cudaGetDevice(), clamped into [0, MAX_DEVICES)
int device_sms();              // cudaDevAttrMultiProcessorCount of the current device (cached per device)

// Opt a kernel in to `bytes` of dynamic shared memory once per device.  `done` is a per-kernel flag array.
template <typename F>
inline int ensure_dyn_smem(F* func, int bytes, bool (&done)[MAX_DEVICES], const char* what) {
    const int dev = current_device();
    if (done[dev]) return 0;
    cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) {
        set_error("%s: cannot opt in to %d B of shared memory on device %d: %s", what, bytes, dev, cudaGetErrorString(e));
        return 1;
    }
    done[dev] = true;
    return 0;
}

// Development knobs (profiling experiments of tools/dev): set explicitly through ft_set_debug(), never read from the
// environment by the library; all zero in production.
enum { FT_DBG_SKIP = 0, FT_DBG_CHUNK = 1, FT_DBG_CORR_PARTS = 2, FT_DBG_CORR_VARIANT = 3, FT_DBG_NO_PDL = 4, FT_DBG_MAX_CTAS = 5, FT_DBG_COUNT = 8 };
int dev_knob(int which);

#define FT_REQUIRE(cond, ...)              \
    do {                                   \
        if (!(cond)) {                     \
            ft::set_error(__VA_ARGS__);    \
            return 2;                      \
        }                                  \
    } while (0)

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

// ---- split-fp16 value helpers -------------------------------------------------------------------
// v = float(hi) + float(lo).  hi is clamped to the finite fp16 range so an out-of-range activation
// saturates instead of turning the whole downstream tensor into inf/NaN.
__device__ __forceinline__ void split_f32(float v, __half& hi, __half& lo) {
    float c = fminf(fmaxf(v, -65504.f), 65504.f);
    hi = __float2half_rn(c);
    lo = __float2half_rn(c - __half2float(hi));
}

// Two values at once: one saturating packed conversion per plane (F2FP.SATFINITE) instead of clamp + two scalar conversions.
// Identical to split_f32 for |v| <= 65504; beyond that hi saturates to +-65504 and lo carries what it can of the rest.
__device__ __forceinline__ void split_f32x2(float v0, float v1, uint32_t& hi2, uint32_t& lo2) {
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi2) : "f"(v1), "f"(v0));
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hi2));
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo2) : "f"(v1 - f.y), "f"(v0 - f.x));
}

__device__ __forceinline__ float join_f16(__half hi, __half lo) { return __half2float(hi) + __half2float(lo); }

struct alignas(16) Half8 {
    __half2 v[4];
};

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
    if (act == FT_ACT_RELU) return fmaxf(v, 0.f);
    if (act == FT_ACT_LEAKY) return v > 0.f ? v : v * slope;
    return v;
}

}  // namespace ft
