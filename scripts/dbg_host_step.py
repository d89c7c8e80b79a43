"""Time r3d_fuser_step_host under a few option sets (debugging the C-ABI e2e number)."""
import sys, time, torch
sys.path.insert(0, '.')
from r3d_b200 import ops, _lib
B, T, C = 64, 512, 512
dt = torch.bfloat16
g = torch.Generator().manual_seed(0)
rgb = torch.randn(B, T, C, generator=g).relu_().to(dt).pin_memory()
dep = torch.randn(B, T, C, generator=g).relu_().to(dt).pin_memory()
gst = torch.randn(B, T, 2, C, generator=g).to(dt).pin_memory()
out = (torch.empty(B, T, 2, C, dtype=dt).pin_memory(), torch.empty(2 * B, dtype=torch.float32).pin_memory(),
       torch.empty(B, T, C, dtype=dt).pin_memory(), torch.empty(B, T, C, dtype=dt).pin_memory(),
       torch.empty(2, C // 4, dtype=torch.int64).pin_memory())
for opts in ({}, {"jacobi_chunks": 1}, {"jacobi_own_streams": 0}, {"jacobi_overlap_v": 0}, {}):
    for k, v in opts.items():
        _lib.set_option(k, v)
    ts = []
    for i in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ops.fuser_step_host(rgb, dep, gst, out=out)
        ts.append((time.perf_counter() - t0) * 1e3)
    print(opts, [round(t, 1) for t in ts], "erank mean", float(out[1].mean()))
    for k in opts:
        _lib.set_option(k, {"jacobi_chunks": 2, "jacobi_own_streams": 1, "jacobi_overlap_v": 1}[k])
