"""Early timings on the GPU box (dev tool): PoseResNet-50 B=64 split/fast vs the oracle module under
cuDNN (fp32 strict / TF32 / bf16), correlation vs the reference kernel."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from flowtrack.pytorch_b200 import _lib  # noqa: E402
from flowtrack.pytorch_b200.pose import models  # noqa: E402
from oracle import pose_oracle  # noqa: E402


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sorted(ts)[len(ts) // 2]


def main():
    B = int(os.environ.get("QB_BATCH", "64"))
    torch.manual_seed(0)
    orc = pose_oracle.deconv("resnet50", 17, False)
    pose_oracle.condition_weights(orc, 0, (128, 96), 4)
    m = models.deconv("resnet50", 17, False)
    m.load_state_dict(orc.state_dict())
    m.cuda()
    x = torch.randn(B, 3, 256, 192, device="cuda")
    gflop = 10.853 * B
    for prec in ("split", "fast"):
        m.precision = prec
        best, med = timeit(lambda: m(x))
        print("ours %-5s B=%d: best %.3f ms median %.3f ms -> %.0f crops/s, %.1f TFLOP/s algorithmic" % (prec, B, best, med, B / best * 1e3, gflop / best), flush=True)
    m.precision = "split"
    _lib.reset_launch_count()
    m.use_cuda_graph = False
    m.refresh()
    m(x)
    print("launches per forward:", _lib.launch_count())
    orc = orc.cuda().eval()
    with torch.no_grad():
        for name, setup in (("fp32-strict", lambda: (setattr(torch.backends.cudnn, "allow_tf32", False), orc.float())),
                            ("tf32", lambda: (setattr(torch.backends.cudnn, "allow_tf32", True), orc.float())),
                            ("bf16-cl", lambda: orc.to(torch.bfloat16).to(memory_format=torch.channels_last))):
            setup()
            xx = x.to(next(orc.parameters()).dtype)
            if name == "bf16-cl":
                xx = xx.to(memory_format=torch.channels_last)
            torch.backends.cudnn.benchmark = True
            best, med = timeit(lambda: orc(xx))
            print("cuDNN %-11s B=%d: best %.3f ms -> %.0f crops/s, %.1f TFLOP/s" % (name, B, best, B / best * 1e3, gflop / best), flush=True)
    # correlation
    a, b = torch.randn(32, 256, 48, 64, device="cuda"), torch.randn(32, 256, 48, 64, device="cuda")
    out = torch.empty(32, 441, 48, 64, device="cuda")
    f = lambda: _lib.check(_lib.lib().ft_correlation_forward(a.data_ptr(), b.data_ptr(), out.data_ptr(), 32, 256, 48, 64, 20, 1, 20, 1, 2, 0, None))
    best, med = timeit(f)
    byts = 32 * 11.71e6
    print("corr ours B=32: best %.3f ms -> %.0f GB/s algorithmic (%.1f%% of 6547)" % (best, byts / best / 1e6, byts / best / 1e6 / 65.47), flush=True)
    so = os.path.join(ROOT, "oracle", "_ref", "libref_correlation.so")
    if os.path.exists(so):
        ref = ctypes.CDLL(so)
        r1 = torch.zeros(32, 88, 104, 256, device="cuda")
        r2 = torch.zeros_like(r1)
        vp, ci = ctypes.c_void_p, ctypes.c_int
        ref.Correlation_forward_cuda_kernel.argtypes = [vp] + [ci] * 8 + [vp] + [ci] * 7 + [vp] + [ci] * 5 + [vp, vp] + [ci] * 6 + [vp]
        g = lambda: ref.Correlation_forward_cuda_kernel(out.data_ptr(), 32, 441, 48, 64, *out.stride(), a.data_ptr(), 256, 48, 64, *a.stride(),
                                                        b.data_ptr(), 256, *b.stride(), r1.data_ptr(), r2.data_ptr(), 20, 1, 20, 1, 2, 1, None)
        best, med = timeit(g, n=5, warm=1)
        print("corr reference kernel B=32: best %.3f ms -> %.0f GB/s algorithmic" % (best, byts / best / 1e6), flush=True)


if __name__ == "__main__":
    main()
