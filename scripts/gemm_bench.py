"""Per-shape timing of the tcgen05 linear GEMM (csrc/linear_tcgen05.cu) at the fuser Block's headline shapes
(rows R = 2 * 64 * 512 = 65536 per modality pair -> 131072 token rows, C = 512, hidden 2048, bf16), next to
torch.matmul (cuBLAS) on the same operands as a library yardstick."""
import os, sys, json
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
wv = torch.randn(C, C, device=dev, dtype=dt) / 22; w1 = torch.randn(Hd, C, device=dev, dtype=dt) / 22; w2 = torch.randn(C, Hd, device=dev, dtype=dt) / 45
b = torch.randn(C, device=dev, dtype=dt); b1 = torch.randn(Hd, device=dev, dtype=dt)
rows = []
def add(name, fn, flops, lib):
    us, ul = t(fn), t(lib)
    rows.append(dict(name=name, us=round(us, 1), tflops=round(flops / us / 1e6, 1), cublas_us=round(ul, 1), cublas_tflops=round(flops / ul / 1e6, 1)))
    print(rows[-1], flush=True)
add("fwd V      (R,C)x(C,C)^T plain", lambda: ops.gemm(x, wv), 2 * R * C * C, lambda: x @ wv.T)
add("fwd proj   +bias +residual", lambda: ops.gemm(x, wv, bias=b, residual=x), 2 * R * C * C, lambda: x @ wv.T)
add("fwd fc1    +bias GELU +aux", lambda: ops.gemm(x, w1, bias=b1, act=1, want_aux=True), 2 * R * C * Hd, lambda: x @ w1.T)
add("fwd fc2    +bias +residual", lambda: ops.gemm(xh, w2, bias=b, residual=x), 2 * R * C * Hd, lambda: xh @ w2.T)
add("bwd dH     dgelu + colsum", lambda: ops.gemm(x, w2, True, False, aux_in=xh, colsum=True), 2 * R * C * Hd, lambda: x @ w2)
add("bwd dh2    (R,4C)x(4C,C)", lambda: ops.gemm(xh, w1, True, False), 2 * R * C * Hd, lambda: xh @ w1)
add("bwd dx     (R,C)x(C,C)", lambda: ops.gemm(x, wv, True, False), 2 * R * C * C, lambda: x @ wv)
add("bwd dW2    (C,R)x(R,4C) splitK", lambda: ops.gemm(x, xh, False, False), 2 * R * C * Hd, lambda: x.T @ xh)
add("bwd dW1    (4C,R)x(R,C) splitK", lambda: ops.gemm(xh, x, False, False), 2 * R * C * Hd, lambda: xh.T @ x)
add("bwd dWv    (C,R)x(R,C) splitK", lambda: ops.gemm(x, x, False, False), 2 * R * C * C, lambda: x.T @ x)
add("colsum     (R,C)", lambda: ops.colsum(x), 0, lambda: x.sum(0))
add("colsum     (R,4C)", lambda: ops.colsum(xh), 0, lambda: xh.sum(0))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/gemm_bench.json", "w"), indent=1)
