/* oracle/ref_shims/THC/THC.h -- TEST INFRASTRUCTURE ONLY.
 * A minimal stand-in for the Torch7 "THC" C API (removed from PyTorch in 1.x) with exactly the entry points that the
 * launchers in the reference's Resample2d_kernel.cu (:188-238) and ChannelNorm_kernel.cu (:82-120) use, so both files
 * compile UNMODIFIED from where they lie under /root/reference into oracle/_ref/libref_warp.so (oracle/Makefile, target
 * `ref`).  A THCudaTensor is a plain {size, stride, data} view filled in by the ctypes caller (oracle/ref_gpu.py); the
 * "state" carries the CUDA stream.  Nothing here allocates. */
#ifndef FT_ORACLE_THC_SHIM_H_
#define FT_ORACLE_THC_SHIM_H_
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

typedef struct THCState { cudaStream_t stream; } THCState;
typedef struct THCudaTensor { long size[4]; long stride[4]; float* data; } THCudaTensor;

static inline long THCudaTensor_nElement(THCState* s, const THCudaTensor* t) { (void)s; return t->size[0] * t->size[1] * t->size[2] * t->size[3]; }
static inline float* THCudaTensor_data(THCState* s, const THCudaTensor* t) { (void)s; return t->data; }
static inline cudaStream_t THCState_getCurrentStream(THCState* s) { return s ? s->stream : 0; }
#define THCudaCheck(err)                                                                   \
    do {                                                                                   \
        cudaError_t e_ = (err);                                                            \
        if (e_ != cudaSuccess) fprintf(stderr, "THCudaCheck: %s\n", cudaGetErrorString(e_)); \
    } while (0)
#endif
