import torch
o=torch.empty(32*441*48*64, device='cuda'); a=torch.empty_like(o)
def t(fn,n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
ms=t(lambda:o.fill_(1.0)); print("fill 186MB: %.1f us %.0f GB/s"%(ms*1e3, o.numel()*4/ms/1e6))
ms=t(lambda:a.copy_(o)); print("copy 186MB: %.1f us %.0f GB/s (r+w)"%(ms*1e3, 2*o.numel()*4/ms/1e6))
big=torch.empty(1<<30, dtype=torch.uint8, device='cuda'); big2=torch.empty_like(big)
ms=t(lambda:big2.copy_(big)); print("copy 1GiB: %.1f us %.0f GB/s (r+w)"%(ms*1e3, 2*big.numel()/ms/1e6))
