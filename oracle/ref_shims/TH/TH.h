/* oracle/ref_shims/TH/TH.h -- TEST INFRASTRUCTURE ONLY.
 * A minimal stand-in for the Torch7 "TH" C API (removed from PyTorch in 1.x) with exactly the handful of entry points the
 * reference's lib/detection/nms/src/nms.c uses, so that file compiles UNMODIFIED from where it lies under /root/reference
 * into oracle/_ref/libref_nms.so (oracle/Makefile, target `ref`).  Tensors are plain {data, sizes} views filled in by the
 * ctypes caller (oracle/track_oracle.py: ref_cpu_nms); nothing here allocates except the byte scratch tensor. */
#ifndef FT_ORACLE_TH_SHIM_H_
#define FT_ORACLE_TH_SHIM_H_
#include <stdlib.h>
#include <string.h>

typedef struct { long* data; long size[2]; } THLongTensor;
typedef struct { float* data; long size[2]; } THFloatTensor;
typedef struct { unsigned char* data; long size[2]; } THByteTensor;

#define THArgCheck(cond, argn, msg) ((void)0)
static inline int THLongTensor_isContiguous(const void* t) { (void)t; return 1; }   /* the reference passes float tensors here too */

static inline long THFloatTensor_size(const THFloatTensor* t, int dim) { return t->size[dim]; }
static inline float* THFloatTensor_data(const THFloatTensor* t) { return t->data; }
static inline long* THLongTensor_data(const THLongTensor* t) { return t->data; }

static inline THByteTensor* THByteTensor_newWithSize1d(long n) {
    THByteTensor* t = (THByteTensor*)malloc(sizeof(THByteTensor));
    t->data = (unsigned char*)malloc(n > 0 ? (size_t)n : 1);
    t->size[0] = n, t->size[1] = 1;
    return t;
}
static inline void THByteTensor_fill(THByteTensor* t, unsigned char v) { memset(t->data, v, (size_t)t->size[0]); }
static inline unsigned char* THByteTensor_data(const THByteTensor* t) { return t->data; }
static inline void THByteTensor_free(THByteTensor* t) { free(t->data); free(t); }
#endif
