"""One launch of each graded kernel for an `ncu --set full` capture: the tcgen05 Gram at a tensor-bound shape and at the
headline shape, and the streaming stages (score, exchange forward/backward) at the headline shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from r3d_b200 import ops, _lib
from r3d_b200.ops import _p, _dt, _stream, check

dev = torch.device("cuda:0")
torch.manual_seed(0)
L = _lib.lib()
for B, T, C in ((16, 4096, 2048), (128, 512, 512)):
    x = torch.randn(B, T, C, device=dev).relu_().to(torch.bfloat16)
    for _ in range(2):
        G = ops.gram(x, ops.GRAM_TCGEN05)
    torch.cuda.synchronize()
    del x, G
B, T, C = 64, 512, 512
rows, k = B * T, C // 4
rgb = torch.randn(B, T, C, device=dev).relu_().to(torch.bfloat16)
dep = torch.randn(B, T, C, device=dev).relu_().to(torch.bfloat16)
g = torch.randn(B, T, 2, C, device=dev).to(torch.bfloat16)
ws = torch.empty(L.r3d_score_workspace_floats(rows, C), dtype=torch.float32, device=dev)
out = torch.empty(B, T, 2, C, dtype=torch.bfloat16, device=dev)
d_r, d_d = torch.empty_like(rgb), torch.empty_like(dep)
idx = torch.stack([torch.randperm(C, device=dev)[:k], torch.randperm(C, device=dev)[:k]]).contiguous()
for _ in range(2):
    check(L.r3d_channel_score_partial(_p(rgb), _p(dep), rows, C, _dt(rgb), _p(ws), _stream()))
    check(L.r3d_exchange_fwd(_p(rgb), _p(dep), _p(idx[0]), _p(idx[1]), k, None, None, 0, _p(out), rows, C, _dt(rgb), _stream()))
    check(L.r3d_exchange_bwd(_p(g), None, None, _p(idx[0]), _p(idx[1]), k, None, None, None, 0, _p(d_r), _p(d_d), None,
                             rows, C, _dt(g), _stream()))
torch.cuda.synchronize()
print("done")
