"""Generate tests/golden/erank_*.npz -- effective-rank fixtures from an implementation INDEPENDENT of the numpy
oracle's formulas: torch float64 ``svdvals`` + entropy written with torch ops, gradient by torch autograd.

    python tests/golden/make_erank_golden.py

The reference repository has no effective-rank code, so these vectors do not pin the oracle to the
reference; they pin oracle/erank_oracle.py (numpy svd + hand-derived gradient) and the CUDA chain to a second,
independently written float64 implementation of SURVEY.md appendix B, including the autograd gradient.

The inputs are regenerated from numpy's PCG64 stream (stable by numpy's compatibility policy), rounded to float32;
`x_sha256` guards against drift.  Stored per fixture: sigma (float64), erank with the 1e-4 numerical-rank cut-off and
with none (appendix B verbatim), the gradient of sum_b erank_b (float32).
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

SHAPES = [(3, 64, 128), (2, 256, 512), (1, 512, 512)]
KINDS = ("relu", "gauss", "decay", "rankdef")


def make_input(kind: str, B: int, T: int, C: int, seed: int) -> np.ndarray:
    """float32 (B, T, C); the four spectra of SURVEY.md 8(d)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, T, C))
    c = np.arange(C)
    if kind == "relu":
        x = np.maximum(x, 0.0) * (1.0 + c / C)
    elif kind == "decay":
        x = x * np.exp(-c / (C / 8.0))
    elif kind == "rankdef":
        n = min(T, C)
        r = max(n // 4, 1)
        a = rng.standard_normal((B, T, r))
        b = rng.standard_normal((B, r, C))
        x = a @ b / np.sqrt(r)
    return x.astype(np.float32)


def erank_t(x: torch.Tensor, rtol: float) -> torch.Tensor:
    s = torch.linalg.svdvals(x)
    if rtol > 0:
        keep = s > rtol * s.amax(dim=-1, keepdim=True)
    else:
        keep = s > 0
    s = torch.where(keep, s, torch.zeros_like(s))
    p = s / s.sum(dim=-1, keepdim=True)
    h = -(torch.where(keep, p * torch.log(torch.where(keep, p, torch.ones_like(p))), torch.zeros_like(p))).sum(-1)
    return torch.exp(h)


def main():
    for kind in KINDS:
        for (B, T, C) in SHAPES:
            seed = 9000 + 97 * KINDS.index(kind) + T + C
            x = make_input(kind, B, T, C, seed)
            xt = torch.from_numpy(x).double().requires_grad_(True)
            er = erank_t(xt, 1e-4)
            (g,) = torch.autograd.grad(er.sum(), xt)
            with torch.no_grad():
                er0 = erank_t(xt, 0.0)
                sigma = torch.linalg.svdvals(xt)
            out = os.path.join(HERE, f"erank_{kind}_T{T}_C{C}.npz")
            np.savez_compressed(out, kind=kind, B=B, T=T, C=C, seed=seed,
                                x_sha256=hashlib.sha256(x.tobytes()).hexdigest(),
                                sigma=sigma.numpy(), erank=er.detach().numpy(), erank_rtol0=er0.numpy(),
                                grad=g.numpy().astype(np.float32))
            print(os.path.basename(out), "erank", er.detach().numpy().round(3), "rtol0", er0.numpy().round(3),
                  os.path.getsize(out) // 1024, "KB")


if __name__ == "__main__":
    main()
