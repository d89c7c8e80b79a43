"""BASELINE.json configs[3]: effective-rank kernel sweep (tokens 64-4096 x channels 128-2048, Gram+Jacobi+entropy).
Reports per-shape time, samples/s, Jacobi sweeps, and accuracy vs the float64 oracle on a few samples."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from r3d_b200 import ops, _lib
from oracle import erank_oracle as EO
dev = torch.device('cuda')
shapes = [(256, 64, 128), (256, 128, 512), (256, 256, 256), (128, 512, 512), (64, 1024, 512), (256, 4096, 128),
          (64, 64, 2048), (32, 1024, 1024), (8, 2048, 2048), (16, 4096, 2048)]
rows = []
for B, T, C in shapes:
    g = torch.Generator(device=dev).manual_seed(T * 7 + C)
    decay = torch.exp(-torch.arange(C, device=dev, dtype=torch.float32) / (C / 8))
    x = (torch.randn(B, T, C, generator=g, device=dev) * decay).to(torch.bfloat16)
    for _ in range(2):
        er, sigma, sw = ops.erank(x, return_aux=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    er, sigma, sw = ops.erank(x, return_aux=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    _lib.profile_enable(True); _lib.profile_read(reset=True)
    ops.erank(x)
    prof = _lib.profile_read(reset=True); _lib.profile_enable(False)
    gram_ms = prof.get("gram", {}).get("ms", 0.0)      # two-pass solver: X X^T (tcgen05 bf16) + Y Y^T (bf16-plane GEMM)
    jac_ms = sum(prof.get(k, {}).get("ms", 0.0) for k in ("jacobi_init", "jacobi_inner", "jacobi_update", "jacobi_extract"))
    nchk = min(B, 2)
    ref = EO.erank(x[:nchk].float().cpu().numpy())
    rel = float(np.abs(er[:nchk].cpu().numpy() - ref).max() / ref.max())
    rows.append(dict(B=B, T=T, C=C, n=min(T, C), ms=ms, samples_per_s=B / ms * 1e3, sweeps=float(sw.float().mean()),
                     erank_mean=float(er.mean()), rel_err_vs_f64=rel, gram_ms=gram_ms, jacobi_ms=jac_ms,
                     jacobi_over_gram=jac_ms / max(gram_ms, 1e-9)))
    print(json.dumps(rows[-1]))
    del x, er, sigma, sw
    torch.cuda.empty_cache()
json.dump(rows, open('gpurun_out/erank_sweep.json', 'w'), indent=1)
