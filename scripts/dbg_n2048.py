"""n = 2048 square decay spectrum: sweeps of the two passes and gradient error for a few floor / cap settings."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from r3d_b200 import ops, _lib
from oracle import erank_oracle as EO
dev = torch.device('cuda')
B, T, C = 4, 2048, 2048
g = torch.Generator(device=dev).manual_seed(T * 7 + C)
decay = torch.exp(-torch.arange(C, device=dev, dtype=torch.float32) / (C / 8))
x = (torch.randn(B, T, C, generator=g, device=dev) * decay).to(torch.bfloat16)
x0 = x[:1].float().cpu().numpy()
ref, gref = EO.erank(x0), EO.erank_bwd(x0, np.ones(1))
for opts in ({}, {"jacobi_nu_pass1": 2048}, {"erank_pass2_sweeps": 10}, {"jacobi_nu_pass1": 2048, "erank_pass2_sweeps": 10}):
    for k, v in opts.items():
        _lib.set_option(k, v)
    xt = x.clone().float().requires_grad_(True)
    er, sg, sw = ops.erank(xt, return_aux=True)
    er.sum().backward()
    ge = np.abs(xt.grad[0].cpu().numpy() - gref[0]).max() / np.abs(gref).max()
    print(opts, "sweeps", sw.tolist(), "erank err %.1e" % (abs(er[0].item() - ref[0]) / ref[0]), "grad err %.1e" % ge, flush=True)
    for k in opts:
        _lib.set_option(k, {"jacobi_nu_pass1": 8192, "erank_pass2_sweeps": 6}[k])
