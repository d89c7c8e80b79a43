"""Small end-to-end pass over every kernel family for compute-sanitizer (memcheck/racecheck)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import r3d_b200
from r3d_b200 import ops
dev = torch.device('cuda')
g = torch.Generator().manual_seed(0)
for dtype in (torch.float32, torch.bfloat16):
    for (B, T, C) in ((2, 9, 37), (2, 40, 128)):
        rgb = torch.relu(torch.randn(B, T, C, generator=g)).to(dtype).to(dev).requires_grad_(True)
        dep = torch.relu(torch.randn(B, T, C, generator=g)).to(dtype).to(dev).requires_grad_(True)
        for variant in ("tokenfusion", "vary", "batchnorm"):
            f = r3d_b200.CMFuser(C, num_heads=1, variant=variant).to(dev).to(dtype).train()
            f.embd_drop.p = 0.0
            st = f.token_fusion(rgb, dep, "test")
            st.float().sum().backward()
for (B, T, C) in ((2, 40, 72), (1, 130, 128), (2, 128, 256)):
    x = torch.randn(B, T, C, generator=g).to(dev).requires_grad_(True)
    er = ops.erank(x)
    er.sum().backward()
xb = torch.randn(1, 256, 256, generator=g).to(torch.bfloat16).to(dev)      # tcgen05 Gram + tensor-core panel update
er, sigma, U, Y, sw = ops._erank_fwd_raw(xb, 1e-4, ops.GRAM_TCGEN05)
ops.token_informativeness(sigma, U)
torch.cuda.synchronize()
print("sanitize pass done", float(er[0]))
