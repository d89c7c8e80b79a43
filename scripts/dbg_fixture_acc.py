"""Gradient / erank error of the CUDA chain on the erank golden fixtures under a few solver options."""
import glob, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_erank_golden import make_input
from r3d_b200 import ops, _lib
dev = torch.device("cuda")
opts = [{}, {"jacobi_nu_pass1": 512}, {"jacobi_nu_pass1": 128}, {"jacobi_nu_pass1": 32},
        {"jacobi_nu_pass2": 0.25}, {"jacobi_nu_pass1": 512, "jacobi_nu_pass2": 0.25},
        {"jacobi_nu_pass1": 128, "jacobi_nu_pass2": 0.25}, {"jacobi_nu_pass1": 8192, "jacobi_nu_pass2": 0.25, "erank_pass2_sweeps": 6}]
names = sys.argv[1:] or ["relu_T512_C512", "gauss_T512_C512", "decay_T512_C512", "rankdef_T512_C512", "relu_T256_C512"]
import time
for nm in names:
    z = np.load(os.path.join(ROOT, "tests", "golden", f"erank_{nm}.npz"))
    x = make_input(str(z["kind"]), int(z["B"]), int(z["T"]), int(z["C"]), int(z["seed"]))
    for o in opts:
        for k, v in o.items():
            _lib.set_option(k, v)
        xt = torch.from_numpy(x).to(dev).requires_grad_(True)
        er, sigma, sw = ops.erank(xt, return_aux=True)
        er.sum().backward()
        g = xt.grad.cpu().numpy()
        ge = np.abs(g - z["grad"]).max() / np.abs(z["grad"]).max()
        ee = np.abs(er.detach().cpu().numpy() - z["erank"]).max() / z["erank"].max()
        print(f"{nm:22s} {str(o):32s} erank {ee:.2e} grad {ge:.2e} sweeps {sw.cpu().numpy()}", flush=True)
        for k in o:
            _lib.set_option(k, {"panel_sym": 1, "jacobi_schedule": 0, "erank_pass2_sweeps": 4, "jacobi_nu_pass1": 2048, "jacobi_nu_pass2": 4}[k])
