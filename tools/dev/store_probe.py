"""Drive build/store_probe.so (tools/dev/store_probe.cu): microseconds to write 186 MB with different store shapes and CTA counts."""
import ctypes
import os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
L = ctypes.CDLL(os.path.join(ROOT, "build", "store_probe.so"))
L.store_probe.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
n = 32 * 441 * 48 * 64
o = torch.empty(n, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for blocks, threads in ((148, 128), (148, 1024), (148 * 8, 256)):
    for mode in (0, 1, 2, 4):
        ts = []
        for _ in range(6):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = L.store_probe(o.data_ptr(), n, mode, blocks, threads, torch.cuda.current_stream().cuda_stream)
            e1.record()
            torch.cuda.synchronize()
            assert rc == 0
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts[1:])[len(ts[1:]) // 2]
        print("blocks %4d x %4d threads  mode %d : %7.1f us  %6.0f GB/s" % (blocks, threads, mode, t * 1e3, n * 4 / t / 1e6), flush=True)
