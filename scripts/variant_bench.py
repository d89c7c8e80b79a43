"""token_fusion forward+backward of the three exchanging variants at the headline shape,
through the nn.Module (the call a FUTR model makes), next to the reference's own op sequence run eagerly on the same GPU
(oracle/torch_port.py).  Rotating input sets (> L2), 20 iterations between two CUDA events.
Usage: python scripts/variant_bench.py [out.json]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import r3d_b200
from oracle.torch_port import PortCMFuser

B, T, C = 64, 512, 512
dev = torch.device("cuda:0")
dtype = torch.bfloat16
pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
g = torch.Generator(device=dev).manual_seed(1)
c = torch.arange(C, device=dev, dtype=torch.float32)
sets = [((torch.randn(B, T, C, generator=g, device=dev).relu_() * (1 + c / C)).to(dtype),
         (torch.randn(B, T, C, generator=g, device=dev).relu_() * (2 - c / C)).to(dtype),
         torch.randn(B, T, 2, C, generator=g, device=dev).to(dtype)) for _ in range(4)]
N_el, es = B * T * C, 2
# algorithmic bytes: score 2N + exchange 4N forward; backward 4N (+2N when grad-alpha needs the inputs);
# BatchNorm adds a statistics pass (2N) forward and the BN backward (reads g 2N + x 2N, writes 2N, twice: reduce + apply)
ALG = {"tokenfusion": (6, 4), "vary": (6, 6), "batchnorm": (6, 10)}
rows = []
for variant in ("tokenfusion", "vary", "batchnorm"):
    torch.manual_seed(0)
    ref = PortCMFuser(C, depth=1, num_heads=8, variant=variant).to(dev).to(dtype).train()
    ours = r3d_b200.CMFuser(C, depth=1, num_heads=8, variant=variant).to(dev).to(dtype).train()
    ours.load_state_dict(ref.state_dict())

    def run(mod, i):
        r, d, gs = sets[i % 4]
        r = r.detach().requires_grad_(True)
        d = d.detach().requires_grad_(True)
        mod.token_fusion(r, d, "test").backward(gs)

    res = {"variant": variant}
    for name, mod in (("ours", ours), ("reference_ops_eager_gpu", ref)):
        for i in range(3):
            run(mod, i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for i in range(n):
            run(mod, i)
        e1.record()
        torch.cuda.synchronize()
        res[name + "_us"] = e0.elapsed_time(e1) / n * 1e3
    fa, ba = ALG[variant]
    nbytes = (fa + ba) * N_el * es
    res["algorithmic_MB"] = nbytes / 1e6
    res["ours_GBs"] = nbytes / res["ours_us"] / 1e3
    res["ours_frac_hbm"] = res["ours_GBs"] / pk
    res["speedup_vs_reference_ops"] = res["reference_ops_eager_gpu_us"] / res["ours_us"]
    rows.append(res)
    print(res, flush=True)
if len(sys.argv) > 1:
    json.dump(rows, open(sys.argv[1], "w"), indent=1)
