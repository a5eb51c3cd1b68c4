// store_probe.cu -- dev probe: how fast 4 warps per SM can write 186 MB with different global-store shapes (B200 boxes on driver
// 580.178 run the correlation epilogue's 4-byte stores ~7x slower than the 580.159 boxes did; same SASS).
#include <cuda_runtime.h>
#include <stdint.h>
// mode 0: 4 B per lane, a warp writes consecutive 128-byte lines
// mode 1: 4 B per lane, corr-like: a warp writes 21 lines one plane (12288 B) apart, then the next group of 21
// mode 2: 16 B per lane, consecutive 512-byte blocks
// mode 3: like mode 1 but 8 B per lane (two adjacent planes' rows interleaved would need this; here just wider lines)
// mode 4: like mode 1 with st.global.cs
extern "C" __global__ void store_probe_kernel(float* out, long long n_floats, int mode, int warps_total) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const long long lines = n_floats / 32;
    if (mode == 0) {
        for (long long l = warp; l < lines; l += warps_total) out[l * 32 + lane] = 1.f;
    } else if (mode == 2) {
        float4* o4 = reinterpret_cast<float4*>(out);
        for (long long l = warp; l < lines / 4; l += warps_total) o4[l * 32 + lane] = make_float4(1.f, 1.f, 1.f, 1.f);
    } else {
        // planes of 3072 floats (12288 B = 96 lines); groups of 21 planes; a warp takes (group, line-in-plane) items
        const long long planes = n_floats / 3072, groups = planes / 21;
        for (long long it = warp; it < groups * 96; it += warps_total) {
            const long long g = it / 96, ln = it % 96;
            float* p = out + (g * 21) * 3072 + ln * 32 + lane;
#pragma unroll
            for (int t = 0; t < 21; ++t) {
                if (mode == 4) __stcs(p + (long long)t * 3072, 1.f);
                else p[(long long)t * 3072] = 1.f;
            }
        }
    }
}
extern "C" int store_probe(float* out, long long n_floats, int mode, int blocks, int threads, void* stream) {
    store_probe_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(out, n_floats, mode, blocks * threads / 32);
    return (int)cudaGetLastError();
}
