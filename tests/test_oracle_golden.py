"""Pin the oracle: every function in oracle/ that restates pinned reference
behaviour is checked against tests/golden/*.npz, which were produced by the
UNMODIFIED reference (tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from conftest import golden_files
from oracle import fuser_oracle as O
from oracle.torch_port import PortCMFuser

VARIANTS = ("tokenfusion", "vary", "batchnorm", "safuser")


def _load(path):
    z = np.load(path)
    d = {k: z[k] for k in z.files}
    sd = {k[3:]: v for k, v in d.items() if k.startswith("sd/")}
    return d, sd


def _variant_of(path):
    return os.path.basename(path).split("_B")[0]


ALL = golden_files()
EXCH = [p for p in ALL if _variant_of(p) != "safuser"]


def test_golden_present():
    assert len(ALL) == 12, ALL


@pytest.mark.parametrize("path", [p for p in ALL if _variant_of(p) in ("tokenfusion", "vary")], ids=os.path.basename)
def test_score_and_bottomk(path):
    d, sd = _load(path)
    C = int(d["C"])
    for m, x in (("r", d["rgb"]), ("d", d["depth"])):
        s = O.channel_score(x)
        np.testing.assert_allclose(s, d[f"eval/score_{m}"], rtol=2e-6, atol=0)
        idx = O.bottomk(d[f"eval/score_{m}"], C // 4)
        # torch.topk(sorted=True) on tie-free scores: same order, bit-exact
        np.testing.assert_array_equal(idx, d[f"eval/idx_{m}"])
        np.testing.assert_array_equal(O.bottomk(s, C // 4), d[f"eval/idx_{m}"])


@pytest.mark.parametrize("path", EXCH, ids=os.path.basename)
def test_token_fusion_eval(path):
    d, sd = _load(path)
    v = _variant_of(path)
    st, ir, idd = O.token_fusion(v, d["rgb"], d["depth"], "test", sd, bn_training=False, return_indices=True)
    np.testing.assert_array_equal(ir, d["eval/idx_r"])
    np.testing.assert_array_equal(idd, d["eval/idx_d"])
    if v == "tokenfusion":
        np.testing.assert_array_equal(st, d["eval/stacked"])          # pure copies: bit-exact
    else:
        np.testing.assert_allclose(st, d["eval/stacked"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("path", EXCH, ids=os.path.basename)
def test_token_fusion_train_and_backward(path):
    d, sd = _load(path)
    v = _variant_of(path)
    st, ir, idd = O.token_fusion(v, d["rgb"], d["depth"], "test", sd, bn_training=True, return_indices=True)
    np.testing.assert_allclose(st, d["train/stacked"], rtol=1e-5, atol=2e-6)
    blend = {"tokenfusion": O.BLEND_SWAP, "vary": O.BLEND_SCALE, "batchnorm": O.BLEND_CONVEX}[v]
    if v != "batchnorm":
        gr, gd, ga = O.exchange_bwd(d["train/g_stacked"], d["rgb"], d["depth"], ir, idd, sd.get("alpha"), blend)
        np.testing.assert_allclose(gr, d["train/tf_grad_rgb"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(gd, d["train/tf_grad_depth"], rtol=1e-6, atol=1e-7)
        if v == "vary":
            np.testing.assert_allclose(ga.reshape(1, 1, -1), d["train/tf_grad/alpha"], rtol=1e-5, atol=1e-6)
    else:
        # BN front end: exchange backward w.r.t. the normalised tensors, checked via alpha only here;
        # the full BN chain is checked through the torch port below
        rn = O.batchnorm_train(d["rgb"], sd["bn_rgb.weight"], sd["bn_rgb.bias"])
        dn = O.batchnorm_train(d["depth"], sd["bn_depth.weight"], sd["bn_depth.bias"])
        _, _, ga = O.exchange_bwd(d["train/g_stacked"], rn[0], dn[0], ir, idd, sd["alpha"], blend)
        np.testing.assert_allclose(ga.reshape(1, 1, -1), d["train/tf_grad/alpha"], rtol=2e-4, atol=2e-5)
        # running statistics after one training call (momentum 0.1, unbiased variance)
        for m, st_ in (("rgb", rn), ("depth", dn)):
            rm = 0.9 * sd[f"bn_{m}.running_mean"] + 0.1 * st_[1]
            rv = 0.9 * sd[f"bn_{m}.running_var"] + 0.1 * st_[3]
            np.testing.assert_allclose(rm, d[f"train/after_tf/bn_{m}.running_mean"], rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(rv, d[f"train/after_tf/bn_{m}.running_var"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("path", ALL, ids=os.path.basename)
def test_fuser_forward_numpy(path):
    d, sd = _load(path)
    v = _variant_of(path)
    out = O.fuser_forward(v, d["rgb"], d["depth"], "test", sd, int(d["heads"]), bn_training=False)
    if v == "safuser":
        y, attn = out
        np.testing.assert_allclose(attn, d["eval/attn"], rtol=0, atol=0)   # exactly {0, 1}
    else:
        y = out
    np.testing.assert_allclose(y, d["eval/y"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("path", ALL, ids=os.path.basename)
def test_torch_port_forward_backward(path):
    d, sd = _load(path)
    v = _variant_of(path)
    C = int(d["C"])
    f = PortCMFuser(C, depth=1, num_heads=int(d["heads"]), variant=v)
    missing = f.load_state_dict({k: torch.from_numpy(x) for k, x in sd.items()}, strict=True)
    rgb, dep = torch.from_numpy(d["rgb"]), torch.from_numpy(d["depth"])
    f.eval()
    with torch.no_grad():
        y = f({"rgb": rgb, "depth": dep}, "test")
    y = y[0] if v == "safuser" else y
    np.testing.assert_allclose(y.numpy(), d["eval/y"], rtol=1e-5, atol=1e-6)
    f.train()
    f.embd_drop.p = 0.0
    r = rgb.clone().requires_grad_(True)
    q = dep.clone().requires_grad_(True)
    y = f({"rgb": r, "depth": q}, "test")
    y = y[0] if v == "safuser" else y
    y.backward(torch.from_numpy(d["train/g_y"]))
    np.testing.assert_allclose(y.detach().numpy(), d["train/y"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(r.grad.numpy(), d["train/grad_rgb"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(q.grad.numpy(), d["train/grad_depth"], rtol=1e-4, atol=1e-6)
    for n, p in f.named_parameters():
        key = "train/grad/" + n
        if key in d:
            np.testing.assert_allclose(p.grad.numpy(), d[key], rtol=1e-4, atol=2e-6, err_msg=n)
    # W_q / W_k get exactly zero gradient
    gq = d["train/grad/blocks.0.attn.qkv.weight"]
    assert np.all(gq[: 2 * C] == 0.0)


def test_train_branch_is_all_ties():
    """mode='train' (tokenfusion.py:40-45): constant score; the build's rule is
    lowest-index-first, torch's CPU artefact differs (documented deviation)."""
    z = np.load(os.path.join(os.path.dirname(ALL[0]), "topk_ties_torch_cpu.npz"))
    for C in (16, 64, 128, 512):
        k = C // 4
        s = O.train_branch_score(3, 5, C)
        assert np.all(s == s[0])
        np.testing.assert_array_equal(O.bottomk(s, k), np.arange(k))
        ref = np.sort(z[f"allties/C{C}"])
        assert len(set(ref.tolist())) == k            # a valid k-subset, but not the index prefix
    # partial ties: zero channels are every third one; ours = the first k of them by index
    C = 64
    s2 = np.arange(C, dtype=np.float32); s2[::3] = 0
    np.testing.assert_array_equal(O.bottomk(s2, C // 4), np.arange(0, C, 3)[: C // 4])


def test_multi_modality_port_reduces_to_the_pinned_two_modality_port():
    """oracle/torch_port.py:PortCMFuserM (the unpinned M-modality extension) must equal the golden-pinned
    PortCMFuser('tokenfusion') for M = 2, forward and gradients."""
    from oracle.torch_port import PortCMFuserM
    torch.manual_seed(0)
    a = PortCMFuser(32, 1, 4, variant="tokenfusion").eval()
    b = PortCMFuserM(32, 1, 4).eval()
    b.load_state_dict(a.state_dict())
    g = torch.Generator().manual_seed(1)
    r = torch.relu(torch.randn(2, 5, 32, generator=g)).requires_grad_(True)
    d = torch.relu(torch.randn(2, 5, 32, generator=g)).requires_grad_(True)
    ya = a({"rgb": r, "depth": d}, "test")
    ya.sum().backward()
    gr = r.grad.clone(); r.grad = None; d.grad = None
    yb = b({"rgb": r, "depth": d}, "test")
    yb.sum().backward()
    assert torch.allclose(ya, yb, atol=1e-6) and torch.allclose(gr, r.grad, atol=1e-6)
