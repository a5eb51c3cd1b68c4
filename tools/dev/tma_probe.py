"""Does cuTensorMapEncodeTiled accept overlapping strides (stride[1] < box[0] * elemsize)?  (dev probe, GPU box)"""
import torch
from cuda.bindings import driver as drv

torch.zeros(1, device="cuda")
buf = torch.zeros(1 << 20, dtype=torch.float16, device="cuda")
U64, U32 = drv.cuuint64_t, drv.cuuint32_t


def enc(dims, strides_bytes, box, swz=drv.CUtensorMapSwizzle.CU_TENSOR_MAP_SWIZZLE_128B):
    r = drv.cuTensorMapEncodeTiled(drv.CUtensorMapDataType.CU_TENSOR_MAP_DATA_TYPE_FLOAT16, len(dims), buf.data_ptr(),
                                   [U64(d) for d in dims], [U64(s) for s in strides_bytes], [U32(b) for b in box], [U32(1)] * len(dims),
                                   drv.CUtensorMapInterleave.CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                                   drv.CUtensorMapL2promotion.CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   drv.CUtensorMapFloatOOBfill.CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
    return r[0]


print("plain          ", enc([64, 96, 128, 4], [128, 128 * 96, 128 * 96 * 128], [64, 16, 8, 1]))
print("overlap 32B    ", enc([64, 96, 128, 4], [32, 32 * 104, 32 * 104 * 128], [64, 16, 8, 1]))
print("overlap 16B    ", enc([64, 96, 128, 4], [16, 16 * 200, 16 * 200 * 128], [64, 16, 8, 1]))
print("non-multiple d3", enc([64, 96, 3, 4], [128, 128 * 96 * 2, 128 * 96 * 7], [64, 16, 8, 1]))
print("box > extent   ", enc([8, 6, 8, 4], [16, 16 * 6, 16 * 48], [64, 2, 8, 8]))
