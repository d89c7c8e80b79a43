"""cProfile of CMFuser.token_fusion fwd+bwd through Python (host overhead hunt)."""
import cProfile, pstats, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, r3d_b200
B, T, C = 64, 512, 512
dev = torch.device("cuda:0")
m = r3d_b200.CMFuser(C, depth=1, num_heads=8).to(dev).to(torch.bfloat16).train()
r = torch.randn(B, T, C, device=dev).relu_().bfloat16()
d = torch.randn(B, T, C, device=dev).relu_().bfloat16()
g = torch.randn(B, T, 2, C, device=dev).bfloat16()
def run():
    a = r.detach().requires_grad_(True); b = d.detach().requires_grad_(True)
    m.token_fusion(a, b, "test").backward(g)
for _ in range(5): run()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(200): run()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
