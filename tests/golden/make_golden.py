"""Generate tests/golden/*.npz from the UNMODIFIED reference at /root/reference.

Run in the build container only (the reference does not exist on the GPU box):

    python tests/golden/make_golden.py

Two shims are applied from outside the reference tree:
a stub ``matplotlib`` (model/extras/transformer.py:15 imports it) and a wrapper
that neutralises the hard-coded ``.to('cuda')`` on the 2x2 mask
(model/futr_safuser_tokenfusion.py:77).  Dropout is set to p=0 on the instance
for the train-mode gradient vectors (it cannot be bit-reproduced across
devices).  Nothing is written under /root/reference.
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    mp = types.ModuleType("matplotlib")
    pp = types.ModuleType("matplotlib.pyplot")
    mp.pyplot = pp
    sys.modules.setdefault("matplotlib", mp)
    sys.modules.setdefault("matplotlib.pyplot", pp)
    sys.path.insert(0, REF)
    import model.futr_safuser_tokenfusion as TF
    import model.futr_safuser_tokenfusion_vary as VY
    import model.futr_safuser_batchnormalization as BN
    import model.futr_safuser_depth as SA

    class _CpuMask:
        def __init__(self, t):
            self.t = t

        def to(self, *a, **k):
            return self.t

    for mod in (TF, VY, BN, SA):
        orig = mod.CMFuser.generate_cross_attention_mask
        mod.CMFuser.generate_cross_attention_mask = staticmethod(lambda sz, _o=orig: _CpuMask(_o(sz)))
    return {"tokenfusion": TF, "vary": VY, "batchnorm": BN, "safuser": SA}


def synth(B, T, C, seed):
    """post-ReLU half-normal, channel c scaled so scores are tie-free."""
    g = torch.Generator().manual_seed(seed)
    c = torch.arange(C, dtype=torch.float32)
    rgb = torch.relu(torch.randn(B, T, C, generator=g)) * (1 + c / C)
    dep = torch.relu(torch.randn(B, T, C, generator=g)) * (2 - c / C)
    # shuffle the channel order so the selected sets are not index prefixes/suffixes
    perm_r = torch.randperm(C, generator=g)
    perm_d = torch.randperm(C, generator=g)
    return rgb[:, :, perm_r].contiguous(), dep[:, :, perm_d].contiguous()


def sd_np(m):
    return {k: v.detach().cpu().numpy().copy() for k, v in m.state_dict().items()}


def one_case(mods, variant, B, T, C, heads, seed):
    mod = mods[variant]
    torch.manual_seed(seed)
    f = mod.CMFuser(dim=C, depth=1, num_heads=heads)
    if variant == "batchnorm":
        # make gamma (the BN-variant score) tie-free and the running stats non-trivial
        with torch.no_grad():
            g = torch.Generator().manual_seed(seed + 7)
            f.bn_rgb.weight.copy_(torch.randn(C, generator=g))
            f.bn_depth.weight.copy_(torch.randn(C, generator=g))
            f.bn_rgb.bias.copy_(0.1 * torch.randn(C, generator=g))
            f.bn_depth.bias.copy_(0.1 * torch.randn(C, generator=g))
            f.bn_rgb.running_mean.copy_(0.5 + 0.1 * torch.randn(C, generator=g))
            f.bn_depth.running_mean.copy_(0.5 + 0.1 * torch.randn(C, generator=g))
            f.bn_rgb.running_var.copy_(0.5 + torch.rand(C, generator=g))
            f.bn_depth.running_var.copy_(0.5 + torch.rand(C, generator=g))
    if variant == "vary":
        with torch.no_grad():
            g = torch.Generator().manual_seed(seed + 9)
            f.alpha.copy_(0.5 + torch.rand(1, 1, C, generator=g))
    rgb, dep = synth(B, T, C, 1234 + seed)
    out = {"B": B, "T": T, "C": C, "heads": heads, "rgb": rgb.numpy(), "depth": dep.numpy()}
    for k, v in sd_np(f).items():
        out["sd/" + k] = v

    # ---- eval: token_fusion + forward ------------------------------------
    f.eval()
    with torch.no_grad():
        if variant == "safuser":
            y, attn = f({"rgb": rgb, "depth": dep})
            out["eval/attn"] = attn.numpy()
        else:
            st = f.token_fusion(rgb, dep, "test")
            out["eval/stacked"] = st.numpy()
            y = f({"rgb": rgb, "depth": dep}, "test")
        out["eval/y"] = y.numpy()
    if variant in ("tokenfusion", "vary"):
        k = C // 4
        s_r = rgb.abs().mean(dim=(0, 1), keepdim=True)
        s_d = dep.abs().mean(dim=(0, 1), keepdim=True)
        out["eval/score_r"] = s_r.reshape(-1).numpy()
        out["eval/score_d"] = s_d.reshape(-1).numpy()
        out["eval/idx_r"] = torch.topk(s_r, k, dim=-1, largest=False)[1].reshape(-1).numpy()
        out["eval/idx_d"] = torch.topk(s_d, k, dim=-1, largest=False)[1].reshape(-1).numpy()
    if variant == "batchnorm":
        k = max(0, int(C * 0.1))
        out["eval/idx_r"] = torch.topk(f.bn_rgb.weight.abs().view(1, 1, C), k, dim=-1, largest=False)[1].reshape(-1).numpy()
        out["eval/idx_d"] = torch.topk(f.bn_depth.weight.abs().view(1, 1, C), k, dim=-1, largest=False)[1].reshape(-1).numpy()

    # ---- train(): gradients through token_fusion alone and through forward --
    f.train()
    f.embd_drop.p = 0.0
    gen = torch.Generator().manual_seed(4321 + seed)
    if variant != "safuser":
        r = rgb.clone().requires_grad_(True)
        d = dep.clone().requires_grad_(True)
        sd_before = {k: v.clone() for k, v in f.state_dict().items()}
        st = f.token_fusion(r, d, "test")
        gst = torch.randn(st.shape, generator=gen)
        params = [p for n, p in f.named_parameters() if n in ("alpha", "bn_rgb.weight", "bn_rgb.bias",
                                                              "bn_depth.weight", "bn_depth.bias")]
        names = [n for n, p in f.named_parameters() if n in ("alpha", "bn_rgb.weight", "bn_rgb.bias",
                                                             "bn_depth.weight", "bn_depth.bias")]
        grads = torch.autograd.grad(st, [r, d] + params, gst, allow_unused=True)
        out["train/stacked"] = st.detach().numpy()
        out["train/g_stacked"] = gst.numpy()
        out["train/tf_grad_rgb"] = grads[0].numpy()
        out["train/tf_grad_depth"] = grads[1].numpy()
        for n, gv in zip(names, grads[2:]):
            if gv is not None:
                out["train/tf_grad/" + n] = gv.numpy()
        if variant == "batchnorm":
            for kk in ("bn_rgb.running_mean", "bn_rgb.running_var", "bn_depth.running_mean", "bn_depth.running_var"):
                out["train/after_tf/" + kk] = f.state_dict()[kk].numpy().copy()
            f.load_state_dict(sd_before)   # undo the running-stat update before the forward pass below

    r = rgb.clone().requires_grad_(True)
    d = dep.clone().requires_grad_(True)
    f.zero_grad()
    if variant == "safuser":
        y, _ = f({"rgb": r, "depth": d})
    else:
        y = f({"rgb": r, "depth": d}, "test")
    gy = torch.randn(y.shape, generator=gen)
    y.backward(gy)
    out["train/y"] = y.detach().numpy()
    out["train/g_y"] = gy.numpy()
    out["train/grad_rgb"] = r.grad.numpy()
    out["train/grad_depth"] = d.grad.numpy()
    for n, p in f.named_parameters():
        if p.grad is not None:
            out["train/grad/" + n] = p.grad.numpy()
    return out


def tie_cases():
    """Document torch.topk(largest=False) on ties: informational only."""
    out = {}
    for C in (16, 64, 128, 512):
        k = C // 4
        s = torch.full((1, 1, C), 1.0 / (3 * 5 * C))
        out[f"allties/C{C}"] = torch.topk(s, k, dim=-1, largest=False)[1].reshape(-1).numpy()
        s2 = torch.arange(C, dtype=torch.float32).view(1, 1, C).clone()
        s2[:, :, ::3] = 0.0
        out[f"thirdzero/C{C}"] = torch.topk(s2, k, dim=-1, largest=False)[1].reshape(-1).numpy()
    return out


def main():
    mods = load_reference()
    cases = [
        # (B, T, C, heads, seed)
        (2, 5, 16, 4, 0),
        (3, 7, 64, 8, 1),
        (2, 6, 40, 4, 2),
    ]
    for variant in ("tokenfusion", "vary", "batchnorm", "safuser"):
        for (B, T, C, heads, seed) in cases:
            data = one_case(mods, variant, B, T, C, heads, seed)
            path = os.path.join(OUT, f"{variant}_B{B}_T{T}_C{C}.npz")
            np.savez_compressed(path, **data)
            print("wrote", path, os.path.getsize(path))
    np.savez_compressed(os.path.join(OUT, "topk_ties_torch_cpu.npz"), **tie_cases())
    print("torch", torch.__version__)


if __name__ == "__main__":
    main()
