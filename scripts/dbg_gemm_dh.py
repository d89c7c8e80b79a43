"""Where does the dH GEMM's time go?  Variants of (R,C)x(C,4C) at the headline shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from r3d_b200 import ops
dev = torch.device("cuda"); dt = torch.bfloat16
R, C, Hd = 131072, 512, 2048
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
x = torch.randn(R, C, device=dev, dtype=dt); xh = torch.randn(R, Hd, device=dev, dtype=dt)
w1 = torch.randn(Hd, C, device=dev, dtype=dt) / 22; w2 = torch.randn(C, Hd, device=dev, dtype=dt) / 45
b1 = torch.randn(Hd, device=dev, dtype=dt)
for name, fn in (("plain  x @ w2 (B MN-major)", lambda: ops.gemm(x, w2, True, False)),
                 ("plain  x @ w1^T (B K-major)", lambda: ops.gemm(x, w1)),
                 ("+dgelu", lambda: ops.gemm(x, w2, True, False, aux_in=xh)),
                 ("+colsum", lambda: ops.gemm(x, w2, True, False, colsum=True)),
                 ("+dgelu +colsum", lambda: ops.gemm(x, w2, True, False, aux_in=xh, colsum=True)),
                 ("fc1: +bias", lambda: ops.gemm(x, w1, bias=b1)),
                 ("fc1: +bias +gelu", lambda: ops.gemm(x, w1, bias=b1, act=1)),
                 ("fc1: +bias +gelu +aux", lambda: ops.gemm(x, w1, bias=b1, act=1, want_aux=True)),
                 ("torch x @ w2", lambda: x @ w2)):
    print(f"{name:32s} {t(fn):8.1f} us", flush=True)
