"""Probe: what does a read-only pass reach on this GPU?  torch reductions over 67 MB .. 2 GiB (bf16)."""
import torch
dev = torch.device("cuda:0")
for mb in (64, 256, 2048):
    n = mb * (1 << 20) // 2
    xs = [torch.randn(n, device=dev, dtype=torch.bfloat16) for _ in range(4 if mb <= 256 else 1)]
    for name, fn in (("sum", lambda x: x.sum(dtype=torch.float32)), ("amax", lambda x: x.amax())):
        for i in range(3):
            fn(xs[i % len(xs)])
        torch.cuda.synchronize()
        N = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(N):
            fn(xs[i % len(xs)])
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / N * 1e3
        print(f"{mb:5d} MB {name:5s} {us:8.1f} us  {mb * 1.048576 / us * 1e3:7.1f} GB/s")
