"""Accuracy sweep: erank fwd/bwd error vs float64 oracle for update impl x Jacobi tolerance."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from r3d_b200 import _lib, ops
from oracle import erank_oracle as EO
dev = torch.device('cuda')
def spectra(kind, B, T, C, seed):
    rng = np.random.default_rng(seed)
    if kind == "relu": return np.maximum(rng.standard_normal((B, T, C)), 0).astype(np.float32)
    if kind == "gauss": return rng.standard_normal((B, T, C)).astype(np.float32)
    if kind == "decay": return (rng.standard_normal((B, T, C)) * np.exp(-np.arange(C) / (C / 8))).astype(np.float32)
cases = [("gauss", 1, 512, 512), ("relu", 1, 512, 512), ("gauss", 2, 128, 128), ("relu", 2, 128, 128), ("decay", 2, 256, 512), ("relu", 2, 256, 512)]
for passes in (1, 2):
    for tol in (1e-6,):        # the library default
        tc = 1
        _lib.set_option("erank_passes", passes); _lib.set_option("jacobi_tol", tol); _lib.set_option("jacobi_max_sweeps", 24)
        out = []
        for kind, B, T, C in cases:
            x = spectra(kind, B, T, C, T * 1000 + C)
            xt = torch.from_numpy(x).to(dev).requires_grad_(True)
            er, sigma, sw = ops.erank(xt, return_aux=True)
            g = np.ones(B, np.float32)
            er.sum().backward()
            ref = EO.erank(x); gref = EO.erank_bwd(x, g)
            e1 = np.abs(er.detach().cpu().numpy() - ref).max() / ref.max()
            e2 = np.abs(xt.grad.cpu().numpy() - gref).max() / np.abs(gref).max()
            out.append(f"{kind}{T}x{C}: er {e1:.1e} grad {e2:.1e} sw {int(sw.max())}")
        print(f"passes={passes} tol={tol:g} | " + " | ".join(out))
