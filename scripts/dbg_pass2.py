import sys, numpy as np, torch
sys.path.insert(0, '.')
from r3d_b200 import _lib, ops
from oracle import erank_oracle as EO
dev = torch.device('cuda')
rng = np.random.default_rng(0)
cases = []
for kind, B, T, C in (("relu", 2, 512, 512), ("gauss", 2, 512, 512), ("relu", 2, 256, 256), ("decay", 2, 256, 512)):
    x = rng.standard_normal((B, T, C)).astype(np.float32)
    if kind == "relu": x = np.maximum(x, 0)
    if kind == "decay": x = x * np.exp(-np.arange(C) / (C / 8)).astype(np.float32)
    cases.append((kind, B, T, C, x, EO.erank(x), EO.erank_bwd(x, np.ones(B, np.float32))))
_lib.set_option("erank_passes", 2)
for cap in (1, 2, 3, 4, 6, 0):
    _lib.set_option("erank_pass2_sweeps", cap)
    out = []
    for kind, B, T, C, x, ref, gref in cases:
        xt = torch.from_numpy(x).to(dev).requires_grad_(True)
        er, sigma, sw = ops.erank(xt, return_aux=True); er.sum().backward()
        e1 = np.abs(er.detach().cpu().numpy() - ref).max() / ref.max()
        e2 = np.abs(xt.grad.cpu().numpy() - gref).max() / np.abs(gref).max()
        out.append(f"{kind}{T}x{C}: er {e1:.1e} grad {e2:.1e} sw2 {int(sw.max())}")
    print(f"pass2 cap={cap} | " + " | ".join(out))
