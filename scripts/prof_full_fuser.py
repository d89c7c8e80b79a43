"""Kernel-time breakdown of the whole CMFuser forward+backward at the headline shape (torch.profiler)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, r3d_b200
from torch.profiler import profile, ProfilerActivity
B, T, C = 64, 512, 512
dev = torch.device("cuda:0"); dt = torch.bfloat16
torch.manual_seed(0)
m = r3d_b200.CMFuser(C, depth=1, num_heads=8).to(dev).to(dt).train()
m.embd_drop.p = 0.0
r = torch.randn(B, T, C, device=dev).relu_().to(dt); d = torch.randn(B, T, C, device=dev).relu_().to(dt)
gy = torch.randn(B, T, C, device=dev, dtype=dt)
def run():
    a = r.detach().requires_grad_(True); b = d.detach().requires_grad_(True)
    m({"rgb": a, "depth": b}, "test").backward(gy)
for _ in range(3): run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5): run()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
