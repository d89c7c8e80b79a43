"""One effective-rank forward at the headline shape (for ncu captures of the Jacobi kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from r3d_b200 import ops, _lib
for kv in os.environ.get("R3D_OPTS", "").split(","):          # e.g. R3D_OPTS=jacobi_chunks=1 (full-batch launches)
    if "=" in kv:
        _lib.set_option(kv.split("=")[0], float(kv.split("=")[1]))

B, T, C = (int(v) for v in (sys.argv[1:4] if len(sys.argv) >= 4 else (128, 512, 512)))
torch.manual_seed(0)
x = torch.randn(B, T, C, device="cuda").relu_().to(torch.bfloat16)
er = ops.erank(x)
torch.cuda.synchronize()
print("erank mean", er.mean().item())
