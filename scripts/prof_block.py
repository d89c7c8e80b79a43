"""One CMFuser forward + backward at the headline shape (B=64, T=512, C=512, bf16): for ncu captures of lin_kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, r3d_b200
B, T, C = 64, 512, 512
dev, dt = torch.device("cuda:0"), torch.bfloat16
torch.manual_seed(0)
m = r3d_b200.CMFuser(C, depth=1, num_heads=8).to(dev).to(dt).train()
m.embd_drop.p = 0.0
r = torch.randn(B, T, C, device=dev).relu_().to(dt).requires_grad_(True)
d = torch.randn(B, T, C, device=dev).relu_().to(dt).requires_grad_(True)
gy = torch.randn(B, T, C, device=dev, dtype=dt)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    m({"rgb": r, "depth": d}, "test").backward(gy)
torch.cuda.synchronize()
print("ok")
