"""oracle/ops_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU oracles for the three native FlowNet2 ops of the reference.  Two forms each:

* ``*_c``      : ctypes call into oracle/_build/libops_oracle.so, the plain-C literal
                 restatement of the reference .cu index math (see ops_oracle.c header).
* ``*_torch``  : vectorised torch restatement (fast; used inside the flow oracle nets).

Reference lines restated:
  correlation  correlation_package/src/correlation_cuda_kernel.cu:10-106, correlation_cuda.c:22-42
  resample2d   resample2d_package/src/Resample2d_kernel.cu:20-66
  channelnorm  channelnorm_package/src/ChannelNorm_kernel.cu:19-51

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.
"""
import ctypes
import math
import os
import subprocess

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_c_oracle():
    """Compile ops_oracle.c (gcc) if the shared object is missing or stale."""
    so = os.path.join(_HERE, "_build", "libops_oracle.so")
    src = os.path.join(_HERE, "ops_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "all"], stdout=subprocess.DEVNULL)
    return so


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build_c_oracle())
        fp = ctypes.POINTER(ctypes.c_float)
        lib.oracle_correlation_forward.argtypes = [fp, fp, fp] + [ctypes.c_int] * 9
        lib.oracle_correlation_forward.restype = ctypes.c_int
        lib.oracle_resample2d_forward.argtypes = [fp, fp, fp] + [ctypes.c_int] * 4
        lib.oracle_resample2d_forward.restype = ctypes.c_int
        lib.oracle_channelnorm_forward.argtypes = [fp, fp] + [ctypes.c_int] * 4
        lib.oracle_channelnorm_forward.restype = ctypes.c_int
        lib.oracle_num_threads.restype = ctypes.c_int
        _LIB = lib
    return _LIB


def c_threads():
    return int(_lib().oracle_num_threads())


def _fptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def correlation_shape(H, W, pad, ksize, max_disp, stride1, stride2):
    border = (ksize - 1) // 2 + max_disp
    d = max_disp // stride2
    outH = int(math.ceil((H + 2 * pad - 2 * border) / float(stride1)))
    outW = int(math.ceil((W + 2 * pad - 2 * border) / float(stride1)))
    return (2 * d + 1) ** 2, outH, outW


def correlation_c(in1, in2, pad=20, ksize=1, max_disp=20, stride1=1, stride2=2):
    a = np.ascontiguousarray(in1.detach().cpu().numpy(), dtype=np.float32)
    b = np.ascontiguousarray(in2.detach().cpu().numpy(), dtype=np.float32)
    B, C, H, W = a.shape
    oc, oh, ow = correlation_shape(H, W, pad, ksize, max_disp, stride1, stride2)
    out = np.zeros((B, oc, oh, ow), dtype=np.float32)
    ok = _lib().oracle_correlation_forward(_fptr(a), _fptr(b), _fptr(out), B, C, H, W,
                                           pad, ksize, max_disp, stride1, stride2)
    assert ok == 1
    return torch.from_numpy(out)


def correlation_torch(in1, in2, pad=20, ksize=1, max_disp=20, stride1=1, stride2=2):
    """Vectorised restatement: zero-pad both maps, slide input2 by (tj,ti)*stride2, multiply, sum over
    the k*k window and channels, divide by k*k*C.  Output channel = (tj+d)*(2d+1) + (ti+d)."""
    B, C, H, W = in1.shape
    oc, oh, ow = correlation_shape(H, W, pad, ksize, max_disp, stride1, stride2)
    krad = (ksize - 1) // 2
    d = max_disp // stride2
    p1 = F.pad(in1, (pad, pad, pad, pad))
    p2 = F.pad(in2, (pad, pad, pad, pad))
    out = in1.new_zeros(B, oc, oh, ow)
    ys = torch.arange(oh) * stride1 + max_disp + krad
    xs = torch.arange(ow) * stride1 + max_disp + krad
    for tj in range(-d, d + 1):
        for ti in range(-d, d + 1):
            acc = in1.new_zeros(B, oh, ow)
            for j in range(-krad, krad + 1):
                for i in range(-krad, krad + 1):
                    a = p1[:, :, (ys + j)][:, :, :, (xs + i)]
                    b = p2[:, :, (ys + tj * stride2 + j)][:, :, :, (xs + ti * stride2 + i)]
                    acc = acc + (a * b).sum(1)
            out[:, (tj + d) * (2 * d + 1) + (ti + d)] = acc / float(ksize * ksize * C)
    return out


def resample2d_c(img, flow):
    a = np.ascontiguousarray(img.detach().cpu().numpy(), dtype=np.float32)
    f = np.ascontiguousarray(flow.detach().cpu().numpy(), dtype=np.float32)
    B, C, H, W = a.shape
    out = np.zeros((f.shape[0], C, f.shape[2], f.shape[3]), dtype=np.float32)
    assert f.shape[2] == H and f.shape[3] == W
    ok = _lib().oracle_resample2d_forward(_fptr(a), _fptr(f), _fptr(out), B, C, H, W)
    assert ok == 1
    return torch.from_numpy(out)


def resample2d_torch(img, flow):
    """Same arithmetic as Resample2d_kernel.cu:38-61 with gathers (clamped corners, un-clamped weights)."""
    B, C, H, W = img.shape
    yy, xx = torch.meshgrid(torch.arange(H, dtype=img.dtype), torch.arange(W, dtype=img.dtype), indexing="ij")
    xf = xx[None] + flow[:, 0]
    yf = yy[None] + flow[:, 1]
    x0 = torch.floor(xf)
    y0 = torch.floor(yf)
    alpha = (xf - x0)[:, None]
    beta = (yf - y0)[:, None]
    xL = x0.long().clamp(0, W - 1)
    xR = (x0 + 1).long().clamp(0, W - 1)
    yT = y0.long().clamp(0, H - 1)
    yB = (y0 + 1).long().clamp(0, H - 1)
    flat = img.reshape(B, C, H * W)

    def g(yi, xi):
        idx = (yi * W + xi).reshape(B, 1, H * W).expand(B, C, H * W)
        return flat.gather(2, idx).reshape(B, C, H, W)

    return ((1 - alpha) * (1 - beta) * g(yT, xL) + alpha * (1 - beta) * g(yT, xR)
            + (1 - alpha) * beta * g(yB, xL) + alpha * beta * g(yB, xR))


def channelnorm_c(x):
    a = np.ascontiguousarray(x.detach().cpu().numpy(), dtype=np.float32)
    B, C, H, W = a.shape
    out = np.zeros((B, 1, H, W), dtype=np.float32)
    ok = _lib().oracle_channelnorm_forward(_fptr(a), _fptr(out), B, C, H, W)
    assert ok == 1
    return torch.from_numpy(out)


def channelnorm_torch(x):
    return torch.sqrt((x * x).sum(1, keepdim=True))
