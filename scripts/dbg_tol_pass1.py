"""Two-pass solver: first-pass significance floor and second-pass sweep cap vs accuracy (fp32 erank and gradient
error against the float64 oracle); the matching step times come from `bench.py --opt key=value`."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from r3d_b200 import _lib, ops
from oracle import erank_oracle as EO
dev = torch.device('cuda')
def spectra(kind, B, T, C, seed):
    rng = np.random.default_rng(seed)
    if kind == "relu": return np.maximum(rng.standard_normal((B, T, C)), 0).astype(np.float32)
    if kind == "gauss": return rng.standard_normal((B, T, C)).astype(np.float32)
    if kind == "rankdef":
        r = min(T, C) // 4
        return (rng.standard_normal((B, T, r)) @ rng.standard_normal((B, r, C))).astype(np.float32)
    if kind == "decay": return (rng.standard_normal((B, T, C)) * np.exp(-np.arange(C) / (C / 8))).astype(np.float32)
cases = [("gauss", 1, 512, 512), ("relu", 1, 512, 512), ("gauss", 2, 128, 128), ("decay", 2, 256, 512), ("relu", 2, 256, 512), ("decay", 1, 512, 512), ("rankdef", 1, 512, 512)]
for tol1, cap in ((1e-6, 6), (1e-6, 2), (1e-6, 1)):
    _lib.set_option("erank_passes", 2); _lib.set_option("jacobi_tol_pass1", tol1); _lib.set_option("jacobi_tol", tol1)
    _lib.set_option("erank_pass2_sweeps", cap)
    out = []
    for kind, B, T, C in cases:
        x = spectra(kind, B, T, C, T * 1000 + C)
        xt = torch.from_numpy(x).to(dev).requires_grad_(True)
        er, sigma, sw = ops.erank(xt, return_aux=True)
        er.sum().backward()
        ref = EO.erank(x); gref = EO.erank_bwd(x, np.ones(B, np.float32))
        e1 = np.abs(er.detach().cpu().numpy() - ref).max() / ref.max()
        e2 = np.abs(xt.grad.cpu().numpy() - gref).max() / np.abs(gref).max()
        out.append(f"{kind}{T}x{C}: er {e1:.1e} grad {e2:.1e} sw {int(sw.max())}")
    print(f"tol={tol1:g} pass2_cap={cap:g} | " + " | ".join(out), flush=True)
