"""Whole CMFuser forward+backward at the headline shape: kernel-time breakdown (torch.profiler), wall time per
iteration of the eager Python path, and the same step replayed from a CUDA graph (no host launch overhead)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, r3d_b200
from torch.profiler import profile, ProfilerActivity
B, T, C = 64, 512, 512
dev = torch.device("cuda:0"); dt = torch.bfloat16
torch.manual_seed(0)
m = r3d_b200.CMFuser(C, depth=1, num_heads=8).to(dev).to(dt).train()
m.embd_drop.p = 0.0
r = torch.randn(B, T, C, device=dev).relu_().to(dt).requires_grad_(True)
d = torch.randn(B, T, C, device=dev).relu_().to(dt).requires_grad_(True)
gy = torch.randn(B, T, C, device=dev, dtype=dt)
def run():
    for p in (r, d): p.grad = None
    m({"rgb": r, "depth": d}, "test").backward(gy)
for _ in range(3): run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5): run()
    torch.cuda.synchronize()
tab = prof.key_averages().table(sort_by="cuda_time_total", row_limit=int(os.environ.get("ROWS", "12")), max_name_column_width=70)
print("\n".join(ln[:72] + ln[128:200] for ln in tab.splitlines()))
N = 20
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(N): run()
torch.cuda.synchronize(); wall = (time.perf_counter() - t0) / N * 1e3
t0 = time.perf_counter()
for _ in range(N): run()
host = (time.perf_counter() - t0) / N * 1e3          # time to ENQUEUE an iteration (host side only)
torch.cuda.synchronize()
print(f"eager: {wall:.3f} ms per iteration (wall), host enqueue {host:.3f} ms")
try:
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): run()
    torch.cuda.current_stream().wait_stream(s)
    for p in m.parameters(): p.grad = None
    with torch.cuda.graph(g):
        run()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(N): g.replay()
    torch.cuda.synchronize()
    print(f"cuda graph replay: {(time.perf_counter() - t0) / N * 1e3:.3f} ms per iteration")
except Exception as ex:
    print("cuda graph capture failed:", repr(ex)[:300])
