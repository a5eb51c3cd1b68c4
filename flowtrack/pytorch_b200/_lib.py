"""ctypes binding of libflowtrack_b200.so (include/flowtrack_b200.h).  No fallback: if the library is
missing or a call fails, this raises."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FT_LIB_PATH") or os.path.join(HERE, "libflowtrack_b200.so")      # FT_LIB_PATH: dev A/B builds (tools/dev)

c_int, c_float, c_void_p, c_ll = ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_longlong


class FtAct(ctypes.Structure):
    _fields_ = [("hi", c_void_p), ("lo", c_void_p), ("n", ctypes.c_int32), ("h", ctypes.c_int32), ("w", ctypes.c_int32),
                ("c", ctypes.c_int32), ("ld", ctypes.c_int32), ("coff", ctypes.c_int32)]


class FtConvDesc(ctypes.Structure):
    _fields_ = [("in_", FtAct), ("out", FtAct), ("res", FtAct), ("out_f32", c_void_p), ("w_tile", c_void_p),
                ("w_gemm", c_void_p), ("scale", c_void_p), ("shift", c_void_p),
                ("kind", ctypes.c_int32), ("kh", ctypes.c_int32), ("kw", ctypes.c_int32), ("stride", ctypes.c_int32),
                ("pad", ctypes.c_int32), ("cout", ctypes.c_int32), ("cout_pad", ctypes.c_int32), ("k_pad", ctypes.c_int32),
                ("block_n", ctypes.c_int32), ("act", ctypes.c_int32), ("slope", c_float), ("precision", ctypes.c_int32),
                ("impl", ctypes.c_int32), ("k_mode", ctypes.c_int32), ("in_xpad_l", ctypes.c_int32), ("in_xpad_r", ctypes.c_int32),
                ("cta_pair", ctypes.c_int32), ("res_blocks", ctypes.c_int32), ("pad_hi", ctypes.c_int32), ("tile_order", ctypes.c_int32),
                ("epilogue_buffers", ctypes.c_int32), ("splitk", ctypes.c_int32), ("splitk_counters_len", ctypes.c_int32),
                ("splitk_ws", c_void_p), ("splitk_ws_bytes", ctypes.c_uint64), ("splitk_counters", c_void_p), ("ctas_per_sm", ctypes.c_int32),
                ("reserved0", ctypes.c_int32)]


ACT_NONE, ACT_RELU, ACT_LEAKY = 0, 1, 2
KIND_CONV, KIND_DECONV = 0, 1
PREC_SPLIT, PREC_FAST = 0, 1
IMPL_TCGEN05, IMPL_SIMT = 0, 1
K_PACKED, K_TAP64, K_ROW = 0, 1, 2
ABI_VERSION = 18        # FT_ABI_VERSION of include/flowtrack_b200.h this binding was written against

# name -> (restype, argtypes): every symbol include/flowtrack_b200.h declares
SIGNATURES = {
    "ft_abi_version": (c_int, []),
    "ft_sizeof": (c_int, [c_int]),
    "ft_last_error": (ctypes.c_char_p, []),
    "ft_launch_count": (c_ll, []),
    "ft_reset_launch_count": (None, []),
    "ft_set_debug": (c_int, [c_int, c_int]),
    "ft_correlation_shape": (c_int, [c_int] * 7 + [ctypes.POINTER(c_int)] * 3),
    "ft_correlation_forward": (c_int, [c_void_p] * 3 + [c_int] * 10 + [c_void_p]),
    "ft_correlation_tc_supported": (c_int, [c_int] * 8),
    "ft_correlation_act": (c_int, [ctypes.POINTER(FtAct)] * 2 + [c_void_p, c_int, c_int, c_void_p]),
    "ft_correlation_act_planar": (c_int, [ctypes.POINTER(FtAct)] * 2 + [c_void_p, c_int, c_int, c_void_p]),
    "ft_pack_corr_planar": (c_int, [c_void_p, ctypes.POINTER(FtAct), c_void_p]),
    "ft_correlation_workspace_bytes": (ctypes.c_size_t, [c_int] * 4),
    "ft_correlation_forward_ws": (c_int, [c_void_p] * 3 + [c_int] * 11 + [c_void_p, ctypes.c_size_t, c_void_p]),
    "ft_resample2d_forward": (c_int, [c_void_p] * 3 + [c_int] * 4 + [c_void_p]),
    "ft_channelnorm_forward": (c_int, [c_void_p] * 2 + [c_int] * 4 + [c_void_p]),
    "ft_conv2d_fprop": (c_int, [ctypes.POINTER(FtConvDesc), c_void_p]),
    "ft_maxpool3x3s2": (c_int, [ctypes.POINTER(FtAct)] * 2 + [c_void_p]),
    "ft_pack_image": (c_int, [c_void_p] + [c_ll] * 4 + [c_int, c_float, ctypes.POINTER(FtAct), c_int, c_int, c_void_p]),
    "ft_pack_image_s2d": (c_int, [c_void_p, c_int, c_int, c_int, ctypes.POINTER(FtAct), c_int, c_int, c_void_p]),
    "ft_pack_pair_ws_floats": (c_int, [c_int]),
    "ft_pack_pair": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p, ctypes.POINTER(FtAct), c_int, c_int, c_void_p]),
    "ft_upsample4x": (c_int, [c_void_p, c_void_p] + [c_int] * 4 + [c_float, c_int, c_void_p]),
    "ft_warp_concat": (c_int, [c_void_p, c_void_p, c_float, c_int, c_int, c_int, ctypes.POINTER(FtAct), c_void_p]),
    "ft_fusion_concat": (c_int, [c_void_p] * 3 + [c_int] * 3 + [ctypes.POINTER(FtAct), c_void_p]),
    "ft_flow_head": (c_int, [c_void_p] * 3 + [c_int] * 3 + [c_void_p]),
    "ft_flow_upsample": (c_int, [c_void_p] * 3 + [c_int] * 3 + [ctypes.POINTER(FtAct), c_void_p]),
    "ft_decode_keypoints": (c_int, [c_void_p] + [c_int] * 4 + [c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p]),
    "ft_pack_pair_u8": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p, ctypes.POINTER(FtAct), c_int, c_int, c_void_p]),
    "ft_track_merge": (c_int, [c_void_p] * 5 + [c_int] * 3 + [c_void_p, c_void_p, c_void_p]),
    "ft_track_select": (c_int, [c_void_p] * 3 + [c_int] * 7 + [c_void_p] * 6),
    "ft_crop_affine": (c_int, [c_void_p, c_int] + [c_ll] * 4 + [c_int] * 3 + [c_void_p] * 3 + [c_int] * 4 + [c_void_p, c_void_p]),
    "ft_crop_affine_s2d": (c_int, [c_void_p, c_int] + [c_ll] * 4 + [c_int] * 3 + [c_void_p] * 3 + [c_int] * 4 + [ctypes.POINTER(FtAct), c_int, c_int, c_void_p]),
    "ft_flip_w": (c_int, [c_void_p, c_void_p, c_ll, c_int, c_void_p]),
    "ft_flip_merge": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ft_nms_oks": (c_int, [c_void_p] * 5 + [c_int, c_int, c_int, c_float, c_float, c_void_p, c_void_p, c_void_p]),
    "ft_nms": (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_int, c_void_p, c_void_p, c_void_p]),
    "ft_flow_to_image": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "ft_box_propagation": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
}

DEBUG_KNOBS = ("FT_DEV_SKIP", "FT_DEV_CHUNK", "FT_DEV_CORR_PARTS", "FT_DEV_CORR_VARIANT", "FT_DEV_NO_PDL", "FT_DEV_MAX_CTAS")     # ft_set_debug(which = index)
_lib = None


def lib():
    """Load the shared library (once).  Raises if it has not been built: there is no other path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libflowtrack_b200.so is not built (%s). Run `python -m flowtrack.pytorch_b200.build`; "
                               "this package has no CPU or PyTorch fallback." % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)        # AttributeError if the symbol is missing
            fn.restype, fn.argtypes = res, args
        if l.ft_abi_version() != ABI_VERSION or l.ft_sizeof(0) != ctypes.sizeof(FtAct) or l.ft_sizeof(1) != ctypes.sizeof(FtConvDesc):
            raise RuntimeError("libflowtrack_b200.so struct layout differs from the ctypes binding (stale build?)")
        _lib = l
        # development knobs of tools/dev: forwarded once, explicitly; the library itself never reads the environment
        for which, name in enumerate(DEBUG_KNOBS):
            if os.environ.get(name):
                l.ft_set_debug(which, int(os.environ[name]))
    return _lib


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError("flowtrack_b200 %s failed (%d): %s" % (what, rc, lib().ft_last_error().decode()))


def launch_count():
    return int(lib().ft_launch_count())


def reset_launch_count():
    lib().ft_reset_launch_count()
