import sys, numpy as np, torch
sys.path.insert(0, '.')
from r3d_b200 import _lib, ops
from oracle import erank_oracle as EO
dev = torch.device('cuda')
rng = np.random.default_rng(0)
for passes in (1, 2):
    _lib.set_option("erank_passes", passes)
    out = []
    for kind, B, T, C in (("relu", 2, 512, 512), ("gauss", 2, 512, 512), ("relu", 2, 256, 256), ("relu", 2, 256, 512)):
        x = rng.standard_normal((B, T, C)).astype(np.float32)
        if kind == "relu": x = np.maximum(x, 0)
        xb = torch.from_numpy(x).to(torch.bfloat16)
        xt = xb.to(dev).requires_grad_(True)
        er = ops.erank(xt); er.sum().backward()
        xf = xb.float().numpy()
        ref = EO.erank(xf); gref = EO.erank_bwd(xf, np.ones(B, np.float32))
        e1 = np.abs(er.detach().cpu().numpy() - ref).max() / ref.max()
        e2 = np.abs(xt.grad.float().cpu().numpy() - gref).max() / np.abs(gref).max()
        out.append(f"{kind}{T}x{C}: er {e1:.1e} grad {e2:.1e}")
    print(f"bf16 passes={passes} | " + " | ".join(out))
