"""Gradient error of the CUDA erank chain vs the float64 oracle for (pass-1 floor, pass-2 floor) pairs on inputs with
a dominant mean component (the case that decides the floors: lambda_max / lambda_bulk large)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from r3d_b200 import ops, _lib
from oracle import erank_oracle as EO
dev = torch.device("cuda")
rng = np.random.default_rng(7)
def mk(kind, T, C):
    x = rng.standard_normal((1, T, C))
    c = np.arange(C)
    if kind == "relu1": x = np.maximum(x, 0) * (1 + c / C)
    if kind == "relu3": x = np.maximum(x, 0) * (1 + 3 * c / C)
    if kind == "shift2": x = x + 2.0
    if kind == "shift8": x = x * (1 + c / C) + 8.0
    if kind == "decay": x = x * np.exp(-c / (C / 8))
    return x.astype(np.float32)
cases = [(k, T, C) for k in ("relu1", "relu3", "shift2", "shift8", "decay") for (T, C) in ((512, 512), (384, 512))]
data = [(k, T, C, mk(k, T, C)) for k, T, C in cases]
refs = [(EO.erank(x), EO.erank_bwd(x, np.ones(1))) for _, _, _, x in data]
_lib.set_option('erank_pass2_sweeps', 6)
for nu1, nu2 in [(float(a), -1e-10) for a in (sys.argv[1:] or ["2048", "4096", "8192"])]:
    _lib.set_option("jacobi_nu_pass1", nu1); _lib.set_option("jacobi_nu_pass2", nu2)
    out = []
    for (k, T, C, x), (er_ref, g_ref) in zip(data, refs):
        xt = torch.from_numpy(x).to(dev).requires_grad_(True)
        er, sg, sw = ops.erank(xt, return_aux=True)
        er.sum().backward()
        ge = np.abs(xt.grad.cpu().numpy() - g_ref).max() / np.abs(g_ref).max()
        out.append(f"{k}{T}:{ge:.1e}/{int(sw[0])}")
    print(f"nu1={nu1} nu2={nu2} | " + " ".join(out), flush=True)
