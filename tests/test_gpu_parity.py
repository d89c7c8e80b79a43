"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the
same seeded inputs, against the committed golden fixtures made from the unmodified
reference, and -- at BASELINE.json's full sizes -- through size-independent
properties.  Bit-exact for indices and pure copies; 1e-4 relative for fp32 values."""
import os

import numpy as np
import pytest
import torch

from conftest import golden_files

pytestmark = pytest.mark.gpu

from oracle import fuser_oracle as O          # noqa: E402
from oracle import erank_oracle as EO         # noqa: E402


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def synth(B, T, C, seed, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    c = torch.arange(C, dtype=torch.float32)
    rgb = torch.relu(torch.randn(B, T, C, generator=g)) * (1 + c / C)
    dep = torch.relu(torch.randn(B, T, C, generator=g)) * (2 - c / C)
    pr, pd = torch.randperm(C, generator=g), torch.randperm(C, generator=g)
    return rgb[:, :, pr].contiguous().to(dtype), dep[:, :, pd].contiguous().to(dtype)


def _variant_of(path):
    return os.path.basename(path).split("_B")[0]


def _load(path):
    z = np.load(path)
    d = {k: z[k] for k in z.files}
    return d, {k[3:]: v for k, v in d.items() if k.startswith("sd/")}


# ------------------------------------------------------------------ golden fixtures
@pytest.mark.parametrize("path", golden_files(), ids=os.path.basename)
def test_golden_forward_backward(path, dev):
    import r3d_b200
    d, sd = _load(path)
    v = _variant_of(path)
    C = int(d["C"])
    f = r3d_b200.CMFuser(C, depth=1, num_heads=int(d["heads"]), variant=v)
    f.load_state_dict({k: torch.from_numpy(x) for k, x in sd.items()}, strict=True)
    f = f.to(dev)
    rgb, dep = torch.from_numpy(d["rgb"]).to(dev), torch.from_numpy(d["depth"]).to(dev)
    f.eval()
    with torch.no_grad():
        if v != "safuser":
            st = f.token_fusion(rgb, dep, "test")
            np.testing.assert_array_equal(f.last_indices[0].cpu().numpy(), d["eval/idx_r"])
            np.testing.assert_array_equal(f.last_indices[1].cpu().numpy(), d["eval/idx_d"])
            if v == "tokenfusion":
                np.testing.assert_array_equal(st.cpu().numpy(), d["eval/stacked"])
            else:
                np.testing.assert_allclose(st.cpu().numpy(), d["eval/stacked"], rtol=1e-5, atol=1e-6)
            y = f({"rgb": rgb, "depth": dep}, "test")
        else:
            y, attn = f({"rgb": rgb, "depth": dep})
            np.testing.assert_array_equal(attn.cpu().numpy(), d["eval/attn"])
    np.testing.assert_allclose(y.cpu().numpy(), d["eval/y"], rtol=1e-4, atol=1e-5)

    # train(): BatchNorm batch statistics, dropout p=0 as in the fixture
    f.train()
    f.embd_drop.p = 0.0
    if v != "safuser":
        sd_before = {k: x.clone() for k, x in f.state_dict().items()}
        r = rgb.clone().requires_grad_(True)
        q = dep.clone().requires_grad_(True)
        st = f.token_fusion(r, q, "test")
        np.testing.assert_allclose(st.detach().cpu().numpy(), d["train/stacked"], rtol=1e-5, atol=2e-6)
        st.backward(torch.from_numpy(d["train/g_stacked"]).to(dev))
        if v == "tokenfusion":
            np.testing.assert_array_equal(r.grad.cpu().numpy(), d["train/tf_grad_rgb"])
            np.testing.assert_array_equal(q.grad.cpu().numpy(), d["train/tf_grad_depth"])
        else:
            np.testing.assert_allclose(r.grad.cpu().numpy(), d["train/tf_grad_rgb"], rtol=1e-4, atol=2e-6)
            np.testing.assert_allclose(q.grad.cpu().numpy(), d["train/tf_grad_depth"], rtol=1e-4, atol=2e-6)
        for n, p in f.named_parameters():
            key = "train/tf_grad/" + n
            if key in d:
                np.testing.assert_allclose(p.grad.cpu().numpy(), d[key], rtol=2e-4, atol=2e-5, err_msg=n)
        if v == "batchnorm":
            for kk in ("bn_rgb.running_mean", "bn_rgb.running_var", "bn_depth.running_mean", "bn_depth.running_var"):
                np.testing.assert_allclose(f.state_dict()[kk].cpu().numpy(), d["train/after_tf/" + kk], rtol=1e-5,
                                           atol=1e-6)
            f.load_state_dict(sd_before)
        f.zero_grad()
    r = rgb.clone().requires_grad_(True)
    q = dep.clone().requires_grad_(True)
    y = f({"rgb": r, "depth": q}, "test")
    y = y[0] if v == "safuser" else y
    np.testing.assert_allclose(y.detach().cpu().numpy(), d["train/y"], rtol=1e-4, atol=1e-5)
    y.backward(torch.from_numpy(d["train/g_y"]).to(dev))
    np.testing.assert_allclose(r.grad.cpu().numpy(), d["train/grad_rgb"], rtol=1e-3, atol=2e-5)
    np.testing.assert_allclose(q.grad.cpu().numpy(), d["train/grad_depth"], rtol=1e-3, atol=2e-5)
    for n, p in f.named_parameters():
        key = "train/grad/" + n
        if key in d and p.grad is not None:
            ref = d[key]
            np.testing.assert_allclose(p.grad.cpu().numpy(), ref, rtol=1e-3, atol=1e-4 * max(1.0, np.abs(ref).max()),
                                       err_msg=n)


# ------------------------------------------------------------------ score / bottom-k / exchange vs oracle
SHAPES = [(1, 1, 4), (3, 7, 64), (2, 5, 37), (8, 256, 512), (2, 33, 1024), (1, 300, 2048), (5, 17, 130)]


@pytest.mark.parametrize("B,T,C", SHAPES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_score_bottomk_exchange_vs_oracle(B, T, C, dtype, dev):
    from r3d_b200 import ops
    rgb, dep = synth(B, T, C, 1234 + C, dtype)
    rn, dn = rgb.float().numpy(), dep.float().numpy()      # oracle in fp32 on the (rounded) inputs
    score = ops.channel_score(rgb.to(dev), dep.to(dev)).cpu().numpy()
    np.testing.assert_allclose(score[0], O.channel_score(rn), rtol=1e-5)
    np.testing.assert_allclose(score[1], O.channel_score(dn), rtol=1e-5)
    k = C // 4
    idx = ops.bottomk(torch.from_numpy(score).to(dev), k).cpu().numpy()
    np.testing.assert_array_equal(idx[0], O.bottomk(score[0], k))      # same scores -> bit-exact indices
    np.testing.assert_array_equal(idx[1], O.bottomk(score[1], k))
    # against indices from the ORACLE's own (float64-accumulated) scores: identical except where two
    # oracle scores are closer than fp32 summation error (a near-tie; documented in DESIGN.md)
    for m, xn in ((0, rn), (1, dn)):
        so = O.channel_score(xn)
        io = O.bottomk(so, k)
        diff = idx[m] != io
        if diff.any():
            assert np.all(np.abs(so[idx[m][diff]] - so[io[diff]]) <= 1e-5 * np.abs(so[io[diff]]))
    ir, idd = torch.from_numpy(idx[0]).to(dev), torch.from_numpy(idx[1]).to(dev)
    out = ops.exchange(rgb.to(dev), dep.to(dev), ir, idd)
    ref = O.exchange_fwd(rn, dn, idx[0], idx[1])
    np.testing.assert_array_equal(out.float().cpu().numpy(), ref)      # pure copies: bit-exact in both dtypes
    # backward (swap): bit-exact
    g = torch.randn(B, T, 2, C, generator=torch.Generator().manual_seed(4321)).to(dtype)
    r = rgb.to(dev).requires_grad_(True)
    q = dep.to(dev).requires_grad_(True)
    ops.exchange(r, q, ir, idd).backward(g.to(dev))
    gr, gd, _ = O.exchange_bwd(g.float().numpy(), rn, dn, idx[0], idx[1])
    if dtype == torch.float32:
        np.testing.assert_array_equal(r.grad.cpu().numpy(), gr)
        np.testing.assert_array_equal(q.grad.cpu().numpy(), gd)
    else:
        np.testing.assert_allclose(r.grad.float().cpu().numpy(), gr, rtol=1e-2, atol=1e-2)
        np.testing.assert_allclose(q.grad.float().cpu().numpy(), gd, rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("blend", [1, 2])
@pytest.mark.parametrize("B,T,C", [(3, 7, 64), (4, 100, 512), (2, 9, 37)])
def test_blend_variants_vs_oracle(blend, B, T, C, dev):
    from r3d_b200 import ops
    rgb, dep = synth(B, T, C, 99 + C)
    g = torch.Generator().manual_seed(5)
    alpha = (0.25 + torch.rand(1, 1, C, generator=g))
    k = C // 4
    idx_r = torch.randperm(C, generator=g)[:k]
    idx_d = torch.randperm(C, generator=g)[:k]
    a = alpha.to(dev).requires_grad_(True)
    r = rgb.to(dev).requires_grad_(True)
    q = dep.to(dev).requires_grad_(True)
    out = ops.exchange(r, q, idx_r.to(dev), idx_d.to(dev), a, blend)
    ref = O.exchange_fwd(rgb.numpy(), dep.numpy(), idx_r.numpy(), idx_d.numpy(), alpha.numpy(), blend)
    np.testing.assert_array_equal(out.detach().cpu().numpy(), ref)   # op-by-op fp32, no FMA contraction
    gs = torch.randn(B, T, 2, C, generator=g)
    out.backward(gs.to(dev))
    gr, gd, ga = O.exchange_bwd(gs.numpy(), rgb.numpy(), dep.numpy(), idx_r.numpy(), idx_d.numpy(), alpha.numpy(), blend)
    np.testing.assert_allclose(r.grad.cpu().numpy(), gr, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(q.grad.cpu().numpy(), gd, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(a.grad.cpu().numpy().reshape(-1), ga, rtol=1e-4, atol=1e-4)


def test_ties_and_edge_cases(dev):
    from r3d_b200 import ops
    # all ties -> index prefix; NaN last; -0 == +0
    s = torch.full((2, 64), 0.125, device=dev)
    np.testing.assert_array_equal(ops.bottomk(s, 16).cpu().numpy(), np.tile(np.arange(16), (2, 1)))
    s = torch.arange(64, dtype=torch.float32, device=dev).flip(0).clone()
    s[3] = float("nan")
    s[10] = -0.0
    s[63] = 0.0
    got = ops.bottomk(s.view(1, -1), 64).cpu().numpy()[0]
    assert got[-1] == 3 and got[0] == 10 and got[1] == 63
    # every third channel dead: ours = first k dead channels by index
    C = 96
    sc = torch.arange(C, dtype=torch.float32, device=dev) + 1
    sc[::3] = 0
    np.testing.assert_array_equal(ops.bottomk(sc.view(1, -1), C // 4).cpu().numpy()[0], np.arange(0, C, 3)[: C // 4])
    # large C through the shared-memory path
    C = 5000
    sc = torch.rand(1, C, generator=torch.Generator().manual_seed(0)).to(dev)
    np.testing.assert_array_equal(ops.bottomk(sc, 1250).cpu().numpy()[0], O.bottomk(sc.cpu().numpy()[0], 1250))
    # k == 0 (BN variant with C < 10) and empty batch
    rgb, dep = synth(2, 3, 8, 0)
    e = torch.empty(0, dtype=torch.int64, device=dev)
    out = ops.exchange(rgb.to(dev), dep.to(dev), e, e)
    np.testing.assert_array_equal(out.cpu().numpy(), np.stack([rgb.numpy(), dep.numpy()], axis=2))
    # CPU tensors are refused (no fallback)
    import r3d_b200
    with pytest.raises(r3d_b200.R3DError):
        ops.channel_score(rgb, dep)
    with pytest.raises(r3d_b200.R3DError):
        ops.exchange(rgb.to(dev), dep.to(dev)[:, :2], e, e)
    with pytest.raises(r3d_b200.R3DError):
        ops.channel_score(rgb.to(dev).double(), dep.to(dev).double())


def test_train_branch_constant_score(dev):
    import r3d_b200
    f = r3d_b200.CMFuser(64, num_heads=4).to(dev)
    rgb, dep = synth(2, 5, 64, 3)
    st = f.token_fusion(rgb.to(dev), dep.to(dev), "train")
    ref = O.token_fusion("tokenfusion", rgb.numpy(), dep.numpy(), "train")
    np.testing.assert_array_equal(st.cpu().numpy(), ref)
    np.testing.assert_array_equal(f.last_indices[0].cpu().numpy(), np.arange(16))


# ------------------------------------------------------------------ full-size properties (BASELINE configs)
@pytest.mark.parametrize("B,T,C,dtype", [(64, 512, 512, torch.bfloat16), (8, 256, 512, torch.float32),
                                         (8, 2048, 1024, torch.bfloat16)])
def test_full_size_properties(B, T, C, dtype, dev):
    from r3d_b200 import ops
    g = torch.Generator(device=dev).manual_seed(7)
    c = torch.arange(C, device=dev, dtype=torch.float32)
    rgb = (torch.relu(torch.randn(B, T, C, generator=g, device=dev)) * (1 + c / C)).to(dtype)
    dep = (torch.relu(torch.randn(B, T, C, generator=g, device=dev)) * (2 - c / C)).to(dtype)
    score = ops.channel_score(rgb, dep)
    ref = torch.stack([rgb.float().abs().mean(dim=(0, 1)), dep.float().abs().mean(dim=(0, 1))])
    assert torch.allclose(score, ref, rtol=1e-4)
    k = C // 4
    idx = ops.bottomk(score, k)
    for m in range(2):
        sel = score[m][idx[m]]
        assert torch.all(sel[1:] >= sel[:-1])                                   # sortedness
        rest = torch.ones(C, dtype=torch.bool, device=dev)
        rest[idx[m]] = False
        assert sel.max() <= score[m][rest].min()                                # it is the bottom-k set
        assert idx[m].unique().numel() == k
    # by construction rgb picks (mostly) low channels and depth high ones: the two sets are disjoint
    assert idx[0].max() < C // 2 <= idx[1].min()
    # bit-exact agreement with torch's own selection on the same scores (tie-free here)
    for m in range(2):
        assert torch.equal(idx[m], torch.topk(score[m], k, largest=False)[1])
    out = ops.exchange(rgb, dep, idx[0], idx[1])
    m_r = torch.zeros(C, dtype=torch.bool, device=dev); m_r[idx[0]] = True
    m_d = torch.zeros(C, dtype=torch.bool, device=dev); m_d[idx[1]] = True
    assert torch.equal(out[:, :, 0], torch.where(m_r, dep, rgb))
    assert torch.equal(out[:, :, 1], torch.where(m_d, rgb, dep))
    # exchanging twice with the same index sets on the exchanged pair restores the inputs (involution)
    # only where the sets are disjoint, which they are here
    back = ops.exchange(out[:, :, 0].contiguous(), out[:, :, 1].contiguous(), idx[0], idx[1])
    # channel in S_r: ex_r = depth, ex_d = depth (unchanged) -> second exchange gives depth again: check the
    # linearity/checksum property instead: per-channel column sums are a permutation-invariant checksum
    cs_in = rgb.float().sum(dim=(0, 1)) + dep.float().sum(dim=(0, 1))
    cs_out = out.float().sum(dim=(0, 1, 2))
    expect = torch.where(m_r, 2 * dep.float().sum(dim=(0, 1)), torch.where(m_d, 2 * rgb.float().sum(dim=(0, 1)), cs_in))
    assert torch.allclose(cs_out, expect, rtol=1e-3)
    assert back.shape == out.shape
    # backward: mask-select identity
    gs = torch.randn(B, T, 2, C, generator=g, device=dev).to(dtype)
    r = rgb.clone().requires_grad_(True)
    q = dep.clone().requires_grad_(True)
    ops.exchange(r, q, idx[0], idx[1]).backward(gs)
    zero = torch.zeros((), dtype=dtype, device=dev)
    exp_r = torch.where(m_r, zero, gs[:, :, 0]) + torch.where(m_d, gs[:, :, 1], zero)
    exp_d = torch.where(m_d, zero, gs[:, :, 1]) + torch.where(m_r, gs[:, :, 0], zero)
    assert torch.equal(r.grad, exp_r) and torch.equal(q.grad, exp_d)


# ------------------------------------------------------------------ effective rank
def _spectra(kind, B, T, C, seed):
    rng = np.random.default_rng(seed)
    if kind == "relu":
        return np.maximum(rng.standard_normal((B, T, C)), 0).astype(np.float32)
    if kind == "decay":
        return (rng.standard_normal((B, T, C)) * np.exp(-np.arange(C) / (C / 8))).astype(np.float32)
    if kind == "rankdef":
        r = max(1, min(T, C) // 4)
        return (rng.standard_normal((B, T, r)) @ rng.standard_normal((B, r, C))).astype(np.float32)
    if kind == "gauss":
        return rng.standard_normal((B, T, C)).astype(np.float32)
    raise ValueError(kind)


ER_SHAPES = [(4, 64, 128), (3, 128, 64), (2, 100, 100), (2, 40, 72), (2, 256, 512), (1, 512, 512), (2, 300, 130),
             (3, 1, 16), (2, 16, 1)]


@pytest.mark.parametrize("kind", ["relu", "decay", "rankdef", "gauss"])
@pytest.mark.parametrize("B,T,C", ER_SHAPES)
def test_erank_vs_oracle_fp32(kind, B, T, C, dev):
    from r3d_b200 import ops
    x = _spectra(kind, B, T, C, seed=T * 1000 + C)
    er, sigma, sweeps = ops.erank(torch.from_numpy(x).to(dev), return_aux=True)
    ref = EO.erank(x)
    got = er.cpu().numpy()
    np.testing.assert_allclose(got, ref, rtol=1e-4, err_msg=f"sweeps={sweeps.cpu().numpy()}")
    # singular values themselves (sorted), relative to sigma_max
    s_ref = EO.singular_values(x)
    s_got = np.sort(sigma.cpu().numpy(), axis=-1)[:, ::-1]
    assert np.abs(s_got - s_ref).max() / s_ref.max() < 2e-5


@pytest.mark.parametrize("B,T,C", [(2, 64, 128), (2, 96, 80), (1, 256, 256)])
def test_erank_bf16(B, T, C, dev):
    from r3d_b200 import ops
    x = torch.from_numpy(_spectra("relu", B, T, C, 5)).to(torch.bfloat16)
    er = ops.erank(x.to(dev))
    ref = EO.erank(x.float().numpy())            # oracle on the bf16-rounded inputs
    np.testing.assert_allclose(er.cpu().numpy(), ref, rtol=1e-2)
    np.testing.assert_allclose(er.cpu().numpy(), ref, rtol=1e-4)   # the chain itself is fp32: expect much better


@pytest.mark.parametrize("kind", ["relu", "gauss"])
@pytest.mark.parametrize("B,T,C", [(3, 24, 40), (2, 64, 48), (2, 128, 128), (1, 200, 320)])
def test_erank_backward_vs_oracle(kind, B, T, C, dev):
    from r3d_b200 import ops
    x = _spectra(kind, B, T, C, seed=11 + T)
    g = np.random.default_rng(3).standard_normal(B).astype(np.float32)
    xt = torch.from_numpy(x).to(dev).requires_grad_(True)
    er = ops.erank(xt)
    (er * torch.from_numpy(g).to(dev)).sum().backward()
    ref = EO.erank_bwd(x, g)
    got = xt.grad.cpu().numpy()
    # The gradient puts O(1) weight (through ln p_j) on the smallest kept singular directions.  The default two-pass
    # solver (second Jacobi pass on the graded G2 = Y Y^T) resolves them to relative accuracy: the 1e-4 bar of
    # BASELINE.json's north_star holds for every spectrum here, square hard-edge ones included.
    assert np.abs(got - ref).max() <= 1e-4 * np.abs(ref).max() + 1e-7


@pytest.mark.gpu
@pytest.mark.parametrize("kind,B,T,C", [("relu", 2, 256, 256), ("decay", 2, 128, 320), ("gauss", 2, 100, 52)])
def test_erank_backward_single_pass_option(kind, B, T, C, dev):
    """erank_passes=1 (the faster single-pass solver): erank unchanged to 1e-5, gradients only within the looser
    single-pass bound (they degrade with the conditioning of the sample).  Also covers the SIMT
    second pass (100x52)."""
    from r3d_b200 import ops, _lib
    x = _spectra(kind, B, T, C, seed=5 + T)
    g = np.ones(B, np.float32)
    ref = EO.erank_bwd(x, g)
    out = {}
    try:
        for passes in (1, 2):
            _lib.set_option("erank_passes", passes)
            xt = torch.from_numpy(x).to(dev).requires_grad_(True)
            er = ops.erank(xt)
            er.sum().backward()
            out[passes] = (er.detach().cpu().numpy(), xt.grad.cpu().numpy())
    finally:
        _lib.set_option("erank_passes", 2)
    assert np.abs(out[1][0] - out[2][0]).max() <= 1e-5 * np.abs(out[2][0]).max()
    assert np.abs(out[1][1] - ref).max() <= 2e-2 * np.abs(ref).max() + 1e-7
    assert np.abs(out[2][1] - ref).max() <= 1e-4 * np.abs(ref).max() + 1e-7


@pytest.mark.parametrize("kind,B,T,C", [("relu", 2, 256, 256), ("gauss", 2, 256, 256), ("gauss", 2, 128, 512)])
def test_erank_backward_bf16(kind, B, T, C, dev):
    """north_star: gradients within 1e-2 relative in bf16 (oracle on the bf16-rounded inputs)."""
    from r3d_b200 import ops
    xb = torch.from_numpy(_spectra(kind, B, T, C, seed=21)).to(torch.bfloat16)
    g = np.ones(B, np.float32)
    xt = xb.to(dev).requires_grad_(True)
    ops.erank(xt).sum().backward()
    ref = EO.erank_bwd(xb.float().numpy(), g)
    got = xt.grad.float().cpu().numpy()
    assert np.abs(got - ref).max() <= 1e-2 * np.abs(ref).max()


@pytest.mark.parametrize("B,T,C,dtype", [(3, 128, 256, torch.float32), (2, 256, 256, torch.float32),
                                         (2, 512, 512, torch.bfloat16)])
def test_fuser_step_matches_oracle(B, T, C, dtype, dev):
    """ops.FuserStep is the fused fwd/bwd step bench.py times (erank of both modalities, score -> bottom-k -> exchange,
    exchange backward + d(mean erank)/dX accumulated): every output against the oracle on the same inputs."""
    from r3d_b200 import ops
    rgb, dep = synth(B, T, C, 77, dtype)
    g = torch.randn(B, T, 2, C, generator=torch.Generator().manual_seed(4321)).to(dtype)
    buf = torch.stack([rgb, dep]).to(dev).contiguous()
    step = ops.FuserStep(B, T, C, dtype, dev)
    out, er, dgrad = step(buf, g.to(dev))
    rn, dn, gn = rgb.float().numpy(), dep.float().numpy(), g.float().numpy()
    k = C // 4
    ir, idd = O.bottomk(O.channel_score(rn), k), O.bottomk(O.channel_score(dn), k)
    np.testing.assert_array_equal(np.sort(step.idx[0].cpu().numpy()), np.sort(ir))        # bit-exact selections
    np.testing.assert_array_equal(np.sort(step.idx[1].cpu().numpy()), np.sort(idd))
    np.testing.assert_array_equal(out.float().cpu().numpy(), O.exchange_fwd(rn, dn, ir, idd))
    x2 = np.concatenate([rn, dn])                                                         # (2B, T, C)
    er_ref = EO.erank(x2)
    np.testing.assert_allclose(er.cpu().numpy(), er_ref, rtol=1e-4)
    gr, gd, _ = O.exchange_bwd(gn, rn, dn, ir, idd)
    de = EO.erank_bwd(x2, np.full(2 * B, 1.0 / (2 * B), np.float32))
    ref = np.stack([gr + de[:B], gd + de[B:]])
    got = dgrad.float().cpu().numpy()
    tol = 1e-4 if dtype == torch.float32 else 1e-2
    assert np.abs(got - ref).max() <= tol * np.abs(ref).max()
    # the erank part alone (what the tolerance is really about): subtract the bit-exact exchange gradient
    if dtype == torch.float32:
        part = got - np.stack([gr, gd])
        assert np.abs(part - np.stack([de[:B], de[B:]])).max() <= 1e-4 * np.abs(de).max() + 1e-7 * np.abs(ref).max()


def test_erank_c5_shape(dev):
    """BASELINE.json configs[4] shape (T=2048, C=1024 per sample, bf16): forward and gradient against the oracle."""
    from r3d_b200 import ops
    x = torch.from_numpy(_spectra("relu", 2, 2048, 1024, 3)).to(torch.bfloat16)
    xt = x.to(dev).requires_grad_(True)
    er = ops.erank(xt)
    er.sum().backward()
    xn = x.float().numpy()
    np.testing.assert_allclose(er.detach().cpu().numpy(), EO.erank(xn), rtol=1e-4)
    ref = EO.erank_bwd(xn, np.ones(2, np.float32))
    assert np.abs(xt.grad.float().cpu().numpy() - ref).max() <= 1e-2 * np.abs(ref).max()


def test_gram_and_jacobi_stages(dev):
    from r3d_b200 import ops
    x = _spectra("relu", 3, 96, 160, 1)
    G = ops.gram(torch.from_numpy(x).to(dev), ops.GRAM_SIMT).cpu().numpy()
    Gr = np.einsum("btc,bsc->bts", x.astype(np.float64), x.astype(np.float64))
    assert np.abs(G - Gr).max() / np.abs(Gr).max() < 1e-6
    x2 = _spectra("relu", 2, 200, 72, 2)    # T >= C -> channel side
    G2 = ops.gram(torch.from_numpy(x2).to(dev), ops.GRAM_SIMT).cpu().numpy()
    Gr2 = np.einsum("btc,btd->bcd", x2.astype(np.float64), x2.astype(np.float64))
    assert G2.shape == (2, 72, 72) and np.abs(G2 - Gr2).max() / np.abs(Gr2).max() < 1e-6
    lam, U, sw = ops.jacobi_eigh(torch.from_numpy(G).to(dev))
    lam_ref = np.linalg.eigvalsh(Gr)
    got = np.sort(lam.cpu().numpy(), axis=-1)
    assert np.abs(got - lam_ref).max() / lam_ref.max() < 1e-4    # plain fp32 Jacobi; erank uses the refined sigma
    Un = U.cpu().numpy().astype(np.float64)      # rows are eigenvectors
    for b in range(3):
        assert np.abs(Un[b] @ Un[b].T - np.eye(96)).max() < 1e-3
        R = Un[b] @ Gr[b] @ Un[b].T
        off = R - np.diag(np.diag(R))
        assert np.abs(off).max() / lam_ref.max() < 1e-4
    assert (sw.cpu().numpy() <= 16).all()


def test_token_informativeness(dev):
    from r3d_b200 import ops
    x = _spectra("relu", 2, 48, 96, 4)
    xt = torch.from_numpy(x).to(dev)
    er, sigma, U, Y, sw = ops._erank_fwd_raw(xt, 1e-4, ops.GRAM_SIMT)
    s = ops.token_informativeness(sigma, U).cpu().numpy()
    ref = EO.token_informativeness(x)
    np.testing.assert_allclose(s, ref, rtol=1e-3, atol=1e-6)
    np.testing.assert_allclose(s.sum(-1), 1.0, rtol=1e-4)


def test_host_buffer_entry(dev):
    from r3d_b200 import ops
    rgb, dep = synth(4, 64, 256, 8)
    out, idx = ops.token_fusion_host(rgb.pin_memory(), dep.pin_memory(), 64)
    ref, ir, idd = O.token_fusion("tokenfusion", rgb.numpy(), dep.numpy(), "test", return_indices=True)
    np.testing.assert_array_equal(idx[0].numpy(), ir)
    np.testing.assert_array_equal(idx[1].numpy(), idd)
    np.testing.assert_array_equal(out.numpy(), ref)


@pytest.mark.parametrize("B,T,C", [(2, 512, 512), (1, 600, 256), (3, 2048, 1024), (2, 1000, 768),   # channel side
                                   (2, 256, 512), (3, 128, 1024), (2, 200, 264), (2, 64, 2048),      # token side
                                   (2, 100, 72), (1, 300, 328), (3, 40, 72)])                         # ragged tiles
def test_gram_tcgen05_vs_oracle(B, T, C, dev):
    """bf16 Gram on tcgen05/TMA (both sides, ragged edges) against float64 numpy and against the SIMT kernel."""
    from r3d_b200 import ops
    x = torch.from_numpy(_spectra("relu", B, T, C, 9)).to(torch.bfloat16)
    xd = x.to(dev)
    G = ops.gram(xd, ops.GRAM_TCGEN05).cpu().numpy()
    xf = x.float().numpy().astype(np.float64)
    Gr = xf.transpose(0, 2, 1) @ xf if T >= C else xf @ xf.transpose(0, 2, 1)
    n = min(T, C)
    assert G.shape == (B, n, n)
    err = np.abs(G - Gr).max() / np.abs(Gr).max()
    Gs = ops.gram(xd, ops.GRAM_SIMT).cpu().numpy()
    err_s = np.abs(Gs - Gr).max() / np.abs(Gr).max()
    print(f"tcgen05 gram rel err {err:.2e} (simt {err_s:.2e}) at T={T}")
    assert err < 1e-5, (err, err_s)      # exact bf16 products; fp32 accumulation over T terms in TMEM


@pytest.mark.parametrize("B,T,C", [(2, 256, 512), (2, 512, 256), (1, 200, 264), (2, 264, 200), (1, 512, 512)])
def test_fp32_gram_on_tensor_cores(B, T, C, dev):
    """fp32 inputs: Gram through the bf16-plane tcgen05 GEMM (6 plane products) vs float64 and vs SIMT."""
    from r3d_b200 import ops
    x = _spectra("relu", B, T, C, 21)
    xd = torch.from_numpy(x).to(dev)
    G = ops.gram(xd, ops.GRAM_TCGEN05).cpu().numpy()
    Gs = ops.gram(xd, ops.GRAM_SIMT).cpu().numpy()
    xf = x.astype(np.float64)
    Gr = xf.transpose(0, 2, 1) @ xf if T >= C else xf @ xf.transpose(0, 2, 1)
    scale = np.abs(Gr).max()
    # tensor-core fp32 accumulation in TMEM truncates (round-toward-zero) once per MMA: a systematic
    # -(number of accumulations) * 2^-25 relative offset (~3-5e-6 here), nearly uniform over G and therefore
    # harmless for sigma / sum(sigma); the SIMT kernel rounds to nearest (1e-7).
    assert np.abs(G - Gr).max() / scale < 1e-5, np.abs(G - Gr).max() / scale
    assert np.abs(Gs - Gr).max() / scale < 2e-6


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,T,C", [(2, 128, 256), (2, 256, 128), (1, 256, 256)])
def test_tensor_core_gemms_match_simt(B, T, C, dtype, dev):
    """A/B: refinement + backward GEMMs on tcgen05 (bf16 planes) vs the SIMT kernels, same inputs."""
    from r3d_b200 import ops, _lib
    x = torch.from_numpy(_spectra("relu", B, T, C, 31)).to(dtype).to(dev)
    outs = {}
    for tc in (1, 0):
        _lib.set_option("gemm_tc", tc)
        try:
            xt = x.clone().requires_grad_(True)
            er, sigma, _ = ops.erank(xt, return_aux=True)
            er.sum().backward()
            outs[tc] = (er.detach().float().cpu().numpy(), np.sort(sigma.cpu().numpy(), -1), xt.grad.float().cpu().numpy())
        finally:
            _lib.set_option("gemm_tc", 1)
    np.testing.assert_allclose(outs[1][0], outs[0][0], rtol=2e-6)
    assert np.abs(outs[1][1] - outs[0][1]).max() / outs[0][1].max() < 2e-6
    gmax = np.abs(outs[0][2]).max()
    tol = 1e-4 if dtype == torch.float32 else 1.6e-2          # bf16 output rounding dominates
    if T == C:
        tol = max(tol, 1e-2)     # square hard-edge spectrum: gradient is ill-conditioned (see test_erank_backward_vs_oracle)
    assert np.abs(outs[1][2] - outs[0][2]).max() / gmax < tol


# ------------------------------------------------------------------ several devices in one process (nn.DataParallel pattern)
def test_two_devices_one_process():
    """The reference trains under nn.DataParallel (main_utkinects.py:129): one process, one host thread per device.
    Kernel attributes, tensor maps and the library's side streams are per device; the same thread must also be able
    to use one device after another."""
    import threading
    from r3d_b200 import ops
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    x = _spectra("relu", 2, 128, 256, 9)
    ref = EO.erank(x)
    rgb, dep = synth(2, 64, 256, 5)
    sref = O.channel_score(rgb.numpy())
    # same thread, device after device
    for d in (0, 1, 0):
        dv = torch.device("cuda", d)
        er = ops.erank(torch.from_numpy(x).to(dv))
        np.testing.assert_allclose(er.cpu().numpy(), ref, rtol=1e-4)
        sc = ops.channel_score(rgb.to(dv), dep.to(dv))
        np.testing.assert_allclose(sc[0].cpu().numpy(), sref, rtol=1e-5)
    # one thread per device, concurrently
    out, err = {}, []

    def work(d):
        try:
            dv = torch.device("cuda", d)
            with torch.cuda.device(dv):
                xt = torch.from_numpy(x).to(dv).requires_grad_(True)
                for _ in range(3):
                    er = ops.erank(xt)
                er.sum().backward()
                torch.cuda.synchronize(dv)
                out[d] = (er.detach().cpu().numpy(), xt.grad.cpu().numpy())
        except Exception as ex:   # pragma: no cover
            err.append(repr(ex))

    ts = [threading.Thread(target=work, args=(d,)) for d in (0, 1)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not err, err
    gref = EO.erank_bwd(x, np.ones(2, np.float32))
    for d in (0, 1):
        np.testing.assert_allclose(out[d][0], ref, rtol=1e-4)
        assert np.abs(out[d][1] - gref).max() <= 1e-4 * np.abs(gref).max()


def test_cmfuser_under_dataparallel():
    """nn.DataParallel over two GPUs (main_utkinects.py:129): every replica scores its own shard, so the result equals
    the single-device module applied to each shard (score_scope='local' semantics), forward and backward."""
    import r3d_b200
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    B, T, C = 4, 40, 128
    rgb, dep = synth(B, T, C, 31)
    torch.manual_seed(0)
    fuser = r3d_b200.CMFuser(C, depth=1, num_heads=4, score_scope="local").to("cuda:0").eval()
    dp = torch.nn.DataParallel(fuser, device_ids=[0, 1])
    r = rgb.to("cuda:0").requires_grad_(True)
    d = dep.to("cuda:0").requires_grad_(True)
    y = dp({"rgb": r, "depth": d}, "test")
    gy = torch.randn(B, T, C, generator=torch.Generator().manual_seed(3)).to("cuda:0")
    y.backward(gy)
    r2 = rgb.to("cuda:0").requires_grad_(True)
    d2 = dep.to("cuda:0").requires_grad_(True)
    ys = [fuser({"rgb": r2[i:i + 2], "depth": d2[i:i + 2]}, "test") for i in (0, 2)]
    yref = torch.cat(ys)
    yref.backward(gy)
    assert torch.allclose(y, yref, rtol=1e-4, atol=1e-5)
    assert torch.allclose(r.grad, r2.grad, rtol=1e-4, atol=1e-5) and torch.allclose(d.grad, d2.grad, rtol=1e-4, atol=1e-5)


# ------------------------------------------------------------------ f1, first step: LayerNorm kernels
@pytest.mark.parametrize("rows,C,dtype", [(7, 128, torch.float32), (1000, 512, torch.float32), (333, 40, torch.float32),
                                          (4096, 512, torch.bfloat16), (50, 1024, torch.bfloat16),
                                          (9, 2048, torch.bfloat16), (65, 136, torch.bfloat16)])
def test_layer_norm_vs_torch(rows, C, dtype, dev):
    """ops.layer_norm against F.layer_norm in fp32 on the same (rounded) inputs: forward, dx, dgamma, dbeta."""
    from r3d_b200 import ops
    g = torch.Generator().manual_seed(rows + C)
    x = (torch.randn(rows, C, generator=g) * 2 + 0.5).to(dtype)
    w = (1 + 0.1 * torch.randn(C, generator=g)).to(dtype)
    b = (0.1 * torch.randn(C, generator=g)).to(dtype)
    gy = torch.randn(rows, C, generator=g).to(dtype)
    xr, wr, br = (t.float().clone().requires_grad_(True) for t in (x, w, b))
    yr = torch.nn.functional.layer_norm(xr, (C,), wr, br, 1e-5)
    yr.backward(gy.float())
    xd, wd, bd = (t.detach().clone().to(dev).requires_grad_(True) for t in (x, w, b))
    y = ops.layer_norm(xd, wd, bd, 1e-5)
    y.backward(gy.to(dev))
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    def close(a, ref, scale_tol):
        a = a.float().cpu(); ref = ref.float()
        return (a - ref).abs().max() <= scale_tol * ref.abs().max() + 1e-6
    assert close(y, yr, tol)
    assert close(xd.grad, xr.grad, 4 * tol)
    assert close(wd.grad, wr.grad, 4 * tol)
    assert close(bd.grad, br.grad, 4 * tol)


def test_layer_norm_rejects_cpu_and_bad_width(dev):
    from r3d_b200 import ops
    from r3d_b200._lib import R3DError
    with pytest.raises(R3DError):
        ops.layer_norm(torch.randn(4, 128), torch.ones(128), torch.zeros(128))
    with pytest.raises(R3DError):
        ops.layer_norm(torch.randn(4, 30, device=dev), torch.ones(30, device=dev), torch.zeros(30, device=dev))


@pytest.mark.parametrize("R,C,dtype", [(5, 128, torch.float32), (300, 512, torch.float32), (2048, 512, torch.bfloat16),
                                       (17, 264, torch.bfloat16)])
def test_layer_norm_mean2_and_swap_add_vs_torch(R, C, dtype, dev):
    """The two fused Block kernels against their torch definitions (fp32 on the rounded inputs), forward and backward:
    layer_norm_mean2 = F.layer_norm(x).mean(-2) on (R, 2, C); swap_add = x + p.flip(-2)."""
    from r3d_b200 import ops
    g = torch.Generator().manual_seed(R * 3 + C)
    x = (torch.randn(R, 2, C, generator=g) * 1.5 - 0.3).to(dtype)
    p = torch.randn(R, 2, C, generator=g).to(dtype)
    w = (1 + 0.1 * torch.randn(C, generator=g)).to(dtype)
    b = (0.1 * torch.randn(C, generator=g)).to(dtype)
    gy = torch.randn(R, C, generator=g).to(dtype)
    gs = torch.randn(R, 2, C, generator=g).to(dtype)
    tol = 1e-5 if dtype == torch.float32 else 1e-2

    def close(a, ref, t):
        a = a.float().cpu(); ref = ref.float()
        return (a - ref).abs().max() <= t * ref.abs().max() + 1e-6

    xr, wr, br = (t.float().clone().requires_grad_(True) for t in (x, w, b))
    yr = torch.nn.functional.layer_norm(xr, (C,), wr, br, 1e-5).mean(dim=-2)
    yr.backward(gy.float())
    xd, wd, bd = (t.detach().clone().to(dev).requires_grad_(True) for t in (x, w, b))
    y = ops.layer_norm_mean2(xd, wd, bd, 1e-5)
    y.backward(gy.to(dev))
    assert y.shape == (R, C)
    assert close(y, yr, tol) and close(xd.grad, xr.grad, 4 * tol)
    assert close(wd.grad, wr.grad, 4 * tol) and close(bd.grad, br.grad, 4 * tol)

    x2, p2 = (t.float().clone().requires_grad_(True) for t in (x, p))
    sr = x2 + p2.flip(-2)
    sr.backward(gs.float())
    x3, p3 = (t.detach().clone().to(dev).requires_grad_(True) for t in (x, p))
    s = ops.swap_add(x3, p3)
    s.backward(gs.to(dev))
    if dtype == torch.float32:
        assert torch.equal(s.cpu(), sr.detach()) and torch.equal(x3.grad.cpu(), x2.grad) and torch.equal(p3.grad.cpu(), p2.grad)
    else:
        assert close(s, sr, tol) and torch.equal(x3.grad.float().cpu(), x2.grad) and torch.equal(p3.grad.float().cpu(), p2.grad)


def test_fuser_step_cuda_graph_capture(dev):
    """The fused step makes ~1000 launches on two streams with device-side convergence control and no host
    synchronisation, so it must be capturable in a CUDA graph and replay bit-identically."""
    from r3d_b200 import ops
    B, T, C = 2, 128, 256
    rgb, dep = synth(B, T, C, 13, torch.bfloat16)
    buf = torch.stack([rgb, dep]).to(dev).contiguous()
    g = torch.randn(B, T, 2, C, generator=torch.Generator().manual_seed(2)).to(torch.bfloat16).to(dev)
    step = ops.FuserStep(B, T, C, torch.bfloat16, dev)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            step(buf, g)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    ref = (step.out.clone(), step.er.clone(), step.dgrad.clone(), step.idx.clone())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step(buf, g)
    for t in (step.out, step.er, step.dgrad):
        t.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(step.out, ref[0]) and torch.equal(step.er, ref[1]) and torch.equal(step.dgrad, ref[2])
    assert torch.equal(step.idx, ref[3])
    # new inputs through the same graph: the static buffers are simply overwritten
    rgb2, dep2 = synth(B, T, C, 14, torch.bfloat16)
    buf.copy_(torch.stack([rgb2, dep2]).to(dev))
    graph.replay()
    torch.cuda.synchronize()
    er_graph = step.er.clone()
    step(buf, g)
    torch.cuda.synchronize()
    assert torch.equal(step.er, er_graph)


@pytest.mark.parametrize("opts", [{"panel_merged": 1, "panel_sym": 0}, {"jacobi_inner_regs": 0},
                                  {"jacobi_update_tc": 0}, {"gemm_tc": 0}, {"jacobi_overlap_v": 0},
                                  {"jacobi_chunks": 1}, {"jacobi_chunks": 4}, {"panel_sym": 0}, {"jacobi_schedule": 1},
                                  {"jacobi_schedule": 1, "panel_sym": 0},
                                  {"jacobi_schedule": 1, "jacobi_inner_regs": 0}, {"jacobi_schedule": 0},
                                  {"jacobi_schedule": 0, "jacobi_overlap_v": 0},
                                  {"jacobi_schedule": 0, "jacobi_inner_regs": 0}])
def test_alternative_kernel_paths_agree(opts, dev):
    """Every selectable kernel path (merged panel schedule, shared-memory inner solver, SIMT panel update, SIMT GEMMs,
    no side stream, two chunks, two-pass panel update through H, spread schedule, round-robin schedule) must reproduce the default path's effective rank and gradient."""
    from r3d_b200 import ops, _lib
    defaults = {"panel_merged": 0, "jacobi_inner_regs": 1, "jacobi_update_tc": 1, "gemm_tc": 1, "jacobi_overlap_v": 1,
                "jacobi_chunks": 2, "jacobi_schedule": 2, "panel_sym": 1}
    x = _spectra("relu", 6, 256, 256, 77)
    ref = EO.erank(x)
    gref = EO.erank_bwd(x, np.ones(6, np.float32))
    try:
        for k, v in opts.items():
            _lib.set_option(k, v)
        xt = torch.from_numpy(x).to(dev).requires_grad_(True)
        er = ops.erank(xt)
        er.sum().backward()
        np.testing.assert_allclose(er.detach().cpu().numpy(), ref, rtol=1e-4)
        # the 1e-4 bar is asserted for the default path elsewhere; the alternatives only have to land next to it
        assert np.abs(xt.grad.cpu().numpy() - gref).max() <= 2e-4 * np.abs(gref).max()
    finally:
        for k, v in defaults.items():
            _lib.set_option(k, v)


def test_chunked_batches_are_bit_identical(dev):
    """jacobi_chunks: the batch is cut into chunks that iterate on separate streams (default 2 when every chunk still
    fills the GPU).  Matrices are independent, so the result must not depend on the chunking -- bit for bit."""
    from r3d_b200 import ops, _lib
    x = torch.from_numpy(_spectra("relu", 80, 256, 256, 5)).to(dev)
    outs = []
    try:
        for nch in (1, 2, 4):
            _lib.set_option("jacobi_chunks", nch)
            xt = x.clone().requires_grad_(True)
            er = ops.erank(xt)
            er.sum().backward()
            outs.append((er.detach().clone(), xt.grad.clone()))
    finally:
        _lib.set_option("jacobi_chunks", 2)
    ref = EO.erank(x[:3].cpu().numpy())
    np.testing.assert_allclose(outs[0][0][:3].cpu().numpy(), ref, rtol=1e-4)
    for er, g in outs[1:]:
        assert torch.equal(er, outs[0][0]) and torch.equal(g, outs[0][1])


def test_panel_round_matches_float64(dev):
    """One Jacobi round on the tensor-core panel kernels against float64: G <- Q^T G Q (panel_sym_kernel, one in-place
    pass) and V <- V Q (panel_update_tc_kernel), random 64 x 64 factors, circle-method and XOR pairings; repeated with a
    capped grid so that every CTA walks several tiles (the mbarrier phases carry over from tile to tile)."""
    from r3d_b200 import _lib
    from r3d_b200.ops import _p, _stream
    L = _lib.lib()

    def pair(m, r, t):
        if r < 0:
            mask = -r
            hb = mask.bit_length() - 1
            a = ((t >> hb) << (hb + 1)) | (t & ((1 << hb) - 1))
            return a, a ^ mask
        if m == 2:
            return 0, 1
        x, y = (r, m - 1) if t == 0 else ((r + t) % (m - 1), (r - t + (m - 1)) % (m - 1))
        return min(x, y), max(x, y)

    for (B, npad, rnd) in ((1, 128, 0), (2, 256, 2), (3, 512, 4), (5, 384, 3), (2, 512, -5), (3, 256, -3)):
        rng = np.random.default_rng(abs(rnd) * 100 + npad + B)
        nb, nt = npad // 32, npad // 64
        A = rng.standard_normal((B, npad, npad)).astype(np.float32)
        G = (A + A.transpose(0, 2, 1)) / 2
        V = rng.standard_normal((B, npad, npad)).astype(np.float32)
        Q = rng.standard_normal((B, nt, 64, 64)).astype(np.float32) / 8

        def blocked(M):
            return torch.from_numpy(np.ascontiguousarray(M.reshape(B, npad, npad // 32, 32).transpose(0, 2, 1, 3))).to(dev)

        def plain(t):
            return t.cpu().numpy().reshape(B, npad // 32, npad, 32).transpose(0, 2, 1, 3).reshape(B, npad, npad)

        Qd = torch.from_numpy(np.ascontiguousarray(Q.transpose(0, 1, 3, 2))).to(dev)
        Qf = np.zeros((B, npad, npad))
        for b in range(B):
            for t in range(nt):
                I, J = pair(nb, rnd, t)
                ix = np.concatenate([np.arange(I * 32, I * 32 + 32), np.arange(J * 32, J * 32 + 32)])
                Qf[b][np.ix_(ix, ix)] = Q[b, t]
        Gref = Qf.transpose(0, 2, 1) @ G.astype(np.float64) @ Qf
        Vref = V.astype(np.float64) @ Qf
        outs = []
        for cap in (0, 3):
            Gd, Vd = blocked(G), blocked(V)
            Hd = torch.zeros_like(Gd)
            scratch = torch.zeros(B * 32 + B * nt, dtype=torch.int32, device=dev)
            try:
                _lib.set_option("panel_grid_cap", cap)
                with torch.cuda.device(dev):
                    _lib.check(L.r3d_debug_panel_round(_p(Gd), _p(Hd), _p(Vd), _p(Qd), B, npad, rnd, _p(scratch), _stream()))
            finally:
                _lib.set_option("panel_grid_cap", 0)
            g, v = plain(Gd), plain(Vd)
            assert np.abs(g - Gref).max() <= 1e-5 * np.abs(Gref).max(), (B, npad, rnd, cap)
            assert np.abs(v - Vref).max() <= 1e-5 * np.abs(Vref).max(), (B, npad, rnd, cap)
            outs.append((Gd, Vd))
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


def test_chained_v_update_matches_float64(dev):
    """panel_vchain_kernel (jacobi_schedule = 2): V <- V Q1 Q2 Q3 for three XOR rounds {a, b, a^b} in ONE pass over V,
    intermediate products on chip -- against the float64 product with random (non-orthogonal) 64 x 64 factors."""
    from r3d_b200 import _lib
    from r3d_b200.ops import _p, _stream
    L = _lib.lib()

    def xor_pair(mask, t):
        hb = mask.bit_length() - 1
        a = ((t >> hb) << (hb + 1)) | (t & ((1 << hb) - 1))
        return a, a ^ mask

    for (B, npad, ga, gb) in ((1, 256, 1, 2), (2, 256, 3, 5), (3, 512, 1, 6), (2, 512, 9, 14), (5, 512, 15, 4),
                              (1, 1024, 21, 10)):
        rng = np.random.default_rng(B * 1000 + npad + ga)
        nt = npad // 64
        V = rng.standard_normal((B, npad, npad)).astype(np.float32)
        Q = rng.standard_normal((3, B, nt, 64, 64)).astype(np.float32) / 8
        Vd = torch.from_numpy(np.ascontiguousarray(V.reshape(B, npad, npad // 32, 32).transpose(0, 2, 1, 3))).to(dev)
        Qd = torch.from_numpy(np.ascontiguousarray(Q.transpose(0, 1, 2, 4, 3))).to(dev)
        scratch = torch.zeros(B * 32 + 3 * B * nt, dtype=torch.int32, device=dev)
        # second run with a capped grid: several tiles per CTA (barrier phases across tiles), as at the headline batch
        Vd2 = Vd.clone()
        with torch.cuda.device(dev):
            _lib.check(L.r3d_debug_vchain(_p(Vd), _p(Qd), B, npad, ga, gb, _p(scratch), _stream()))
            try:
                _lib.set_option("panel_grid_cap", 3)
                _lib.check(L.r3d_debug_vchain(_p(Vd2), _p(Qd), B, npad, ga, gb, _p(scratch), _stream()))
            finally:
                _lib.set_option("panel_grid_cap", 0)
        assert torch.equal(Vd, Vd2)
        ref = V.astype(np.float64)
        for k, mask in enumerate((ga, gb, ga ^ gb)):
            Qf = np.zeros((B, npad, npad))
            for b in range(B):
                for t in range(nt):
                    I, J = xor_pair(mask, t)
                    ix = np.concatenate([np.arange(I * 32, I * 32 + 32), np.arange(J * 32, J * 32 + 32)])
                    Qf[b][np.ix_(ix, ix)] = Q[k, b, t]
            ref = ref @ Qf
        got = Vd.cpu().numpy().reshape(B, npad // 32, npad, 32).transpose(0, 2, 1, 3).reshape(B, npad, npad)
        assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max(), (B, npad, ga, gb)


# ------------------------------------------------------------------ round 2 additions
def test_packed_statistic_and_scaled_bottomk(dev):
    """r3d_score_finalize_packed + r3d_bottomk_scaled (no ATen kernels in the step): the packed statistic
    [sum|rgb| | sum|depth| | sum er | rows] and the selection on sums / rows against the oracle, bit-exact indices."""
    from r3d_b200 import ops
    for (B, T, C, dt) in ((3, 17, 64, torch.float32), (8, 256, 512, torch.float32), (4, 64, 200, torch.bfloat16)):
        rgb, dep = synth(B, T, C, 5 + B, dt)
        er = torch.linspace(1.0, 2.0, 2 * B)
        packed = ops.channel_score_packed(rgb.to(dev), dep.to(dev), er.to(dev))
        p = packed.cpu().numpy()
        rn, dn = rgb.float().numpy(), dep.float().numpy()
        ref = np.stack([O.channel_score(rn), O.channel_score(dn)])
        np.testing.assert_allclose(p[:2 * C].reshape(2, C) / (B * T), ref, rtol=1e-5)
        assert p[2 * C + 1] == B * T
        np.testing.assert_allclose(p[2 * C], er.sum().item(), rtol=1e-6)
        k = C // 4
        idx, score = ops.bottomk_packed(packed, k, return_score=True)
        sc = score.cpu().numpy()
        np.testing.assert_array_equal(sc, (p[:2 * C] / p[2 * C + 1]).astype(np.float32).reshape(2, C))   # IEEE division
        for m in range(2):
            np.testing.assert_array_equal(idx[m].cpu().numpy(), O.bottomk(sc[m], k))
            np.testing.assert_array_equal(idx[m].cpu().numpy(), O.bottomk(ref[m], k))   # tie-free synthetic scores


def _erank_fixtures():
    import glob
    from conftest import GOLDEN
    return sorted(glob.glob(os.path.join(GOLDEN, "erank_*.npz")))


@pytest.mark.parametrize("path", _erank_fixtures(), ids=os.path.basename)
def test_erank_golden_fixtures(path, dev):
    """The CUDA chain against tests/golden/erank_*.npz: torch float64 svdvals + autograd, an implementation independent
    of the numpy oracle (tests/golden/make_erank_golden.py).  north_star bars: 1e-4 relative in fp32."""
    import sys
    from conftest import GOLDEN
    sys.path.insert(0, GOLDEN)
    from make_erank_golden import make_input
    from r3d_b200 import ops
    z = np.load(path)
    B, T, C = int(z["B"]), int(z["T"]), int(z["C"])
    x = make_input(str(z["kind"]), B, T, C, int(z["seed"]))
    xt = torch.from_numpy(x).to(dev).requires_grad_(True)
    er, sigma, sweeps = ops.erank(xt, return_aux=True)
    if str(z["kind"]) != "rankdef":
        # (in the noise null space of an exactly rank-deficient sample rotations never die out; the refinement makes
        # the result independent of them, which the value checks below show)
        assert (sweeps > 0).all(), f"eigensolver hit its sweep cap: {sweeps.cpu().numpy()}"
    np.testing.assert_allclose(er.detach().cpu().numpy(), z["erank"], rtol=1e-4)
    sg = np.sort(sigma.cpu().numpy(), axis=1)[:, ::-1]
    assert np.abs(sg - z["sigma"]).max() <= 1e-4 * z["sigma"].max()
    er.sum().backward()
    g = xt.grad.cpu().numpy()
    assert np.abs(g - z["grad"]).max() <= 1e-4 * np.abs(z["grad"]).max()
    # rtol = 0 (appendix B verbatim, no cut-off) is supported too; it agrees with the float64 value wherever the
    # spectrum stays above what an fp32 chain resolves (every fixture but the exactly rank-deficient ones, whose null
    # space comes out as noise of ~1e-6 sigma_max instead of 1e-16)
    if str(z["kind"]) != "rankdef":
        er0 = ops.erank(xt.detach(), rtol=0.0)
        np.testing.assert_allclose(er0.cpu().numpy(), z["erank_rtol0"], rtol=2e-4)


def test_erank_from_short_lived_threads_keeps_stream_sets_flat(dev):
    """nn.DataParallel starts fresh Python threads for every forward (main_utkinects.py:129); the library's side
    streams / events are pooled per device and lent to threads, so their number must not grow with the thread count."""
    import threading
    from r3d_b200 import ops, _lib
    x = torch.from_numpy(_spectra("relu", 2, 128, 128, 11)).to(dev)
    ref = ops.erank(x).cpu()
    torch.cuda.synchronize()
    before = _lib.lib().r3d_stream_sets_created()
    outs = []

    def work():
        with torch.cuda.device(dev):
            outs.append(ops.erank(x).cpu())

    for _ in range(12):
        t = threading.Thread(target=work)
        t.start()
        t.join()
    after = _lib.lib().r3d_stream_sets_created()
    assert after - before <= 1, (before, after)
    for o in outs:
        assert torch.equal(o, ref)


def test_erank_strict_mode_reports_sweep_cap(dev):
    """A capped eigensolver is reported (negative sweeps) and strict mode raises instead of returning silently."""
    from r3d_b200 import ops, _lib
    from r3d_b200._lib import R3DError
    x = torch.from_numpy(_spectra("relu", 2, 256, 256, 3)).to(dev)
    try:
        _lib.set_option("erank_passes", 1)
        _lib.set_option("jacobi_max_sweeps", 2)
        er, sigma, sweeps = ops.erank(x, return_aux=True)
        assert (sweeps < 0).all()
        with pytest.raises(R3DError):
            ops.erank(x, strict=True)
    finally:
        _lib.set_option("erank_passes", 2)
        _lib.set_option("jacobi_max_sweeps", 16)
    er, sigma, sweeps = ops.erank(x, return_aux=True, strict=True)
    assert (sweeps > 0).all()


def test_empty_batch_blend_backward_is_zero(dev):
    """rows == 0: no partial sums are written; d_alpha must be zeros, not workspace garbage."""
    from r3d_b200 import ops
    C = 64
    rgb = torch.zeros(0, 5, C, device=dev, requires_grad=True)
    dep = torch.zeros(0, 5, C, device=dev, requires_grad=True)
    alpha = torch.ones(1, 1, C, device=dev, requires_grad=True)
    idx = torch.arange(16, device=dev)
    out = ops.exchange(rgb, dep, idx, idx, alpha, ops.BLEND_SCALE)
    out.sum().backward()
    assert torch.count_nonzero(alpha.grad) == 0


# ------------------------------------------------------------------ N1: token-axis selection (unpinned; oracle = restatement)
@pytest.mark.parametrize("B,T,C,dt", [(3, 64, 128, torch.float32), (2, 128, 64, torch.float32), (2, 96, 96, torch.float32),
                                      (2, 256, 512, torch.bfloat16), (1, 37, 20, torch.float32)])
def test_token_axis_scores_selection_exchange(B, T, C, dt, dev):
    from r3d_b200 import ops
    rgb, dep = synth(B, T, C, 31 + T, dt)
    # give the tokens distinct weights so that the scores are tie-free and well separated
    w = (1.0 + torch.arange(T, dtype=torch.float32) / T).view(1, T, 1)
    rgb, dep = (rgb.float() * w).to(dt), (dep.float() * w.flip(1)).to(dt)
    rn, dn = rgb.float().numpy(), dep.float().numpy()
    s_r = ops.token_scores(rgb.to(dev)).cpu().numpy()
    s_d = ops.token_scores(dep.to(dev)).cpu().numpy()
    ref_r, ref_d = EO.token_scores(rn), EO.token_scores(dn)
    tol = 2e-3 if dt == torch.bfloat16 else 2e-4
    np.testing.assert_allclose(s_r, ref_r, rtol=tol, atol=tol * ref_r.max())
    np.testing.assert_allclose(s_d, ref_d, rtol=tol, atol=tol * ref_d.max())
    np.testing.assert_allclose(s_r.sum(1), 1.0, rtol=1e-3)
    k = T // 4
    idx_r = ops.bottomk(torch.from_numpy(s_r).to(dev), k)
    idx_d = ops.bottomk(torch.from_numpy(s_d).to(dev), k)
    # selection on the device scores is exactly the oracle rule (ascending score, ties -> lower index) ...
    st_ref, ir, idd = O.token_fusion_tokens(rn, dn, k, scores=(s_r, s_d), return_indices=True)
    np.testing.assert_array_equal(idx_r.cpu().numpy(), ir)
    np.testing.assert_array_equal(idx_d.cpu().numpy(), idd)
    # ... and the SET equals the float64 oracle's wherever the k-th and (k+1)-th scores are separated by more than the
    # score error
    for got, ref in ((ir, ref_r), (idd, ref_d)):
        for b in range(B):
            srt = np.sort(ref[b])
            if srt[k] - srt[k - 1] > 4 * tol * ref.max():
                assert set(got[b].tolist()) == set(np.argsort(ref[b], kind="stable")[:k].tolist())
    r = rgb.to(dev).requires_grad_(True)
    d = dep.to(dev).requires_grad_(True)
    st = ops.token_exchange(r, d, idx_r, idx_d)
    np.testing.assert_array_equal(st.detach().float().cpu().numpy(), st_ref)            # pure copies: bit-exact
    g = torch.randn(B, T, 2, C, generator=torch.Generator().manual_seed(5)).to(dt)
    st.backward(g.to(dev))
    gr, gd = O.token_exchange_bwd(g.float().numpy(), ir, idd)
    rnd = lambda a: torch.from_numpy(a).to(dt).float().numpy()        # a token in both sets adds two terms: one rounding
    np.testing.assert_array_equal(r.grad.float().cpu().numpy(), rnd(gr))
    np.testing.assert_array_equal(d.grad.float().cpu().numpy(), rnd(gd))


def test_token_axis_fuser_module(dev):
    import r3d_b200
    B, T, C = 2, 32, 64
    rgb, dep = synth(B, T, C, 3)
    w = (1.0 + torch.arange(T, dtype=torch.float32) / T).view(1, T, 1)
    rgb, dep = rgb * w, dep * w.flip(1)
    f = r3d_b200.CMFuser(C, depth=1, num_heads=4, select_axis="token").to(dev).eval()
    r = rgb.to(dev).requires_grad_(True)
    d = dep.to(dev).requires_grad_(True)
    y = f({"rgb": r, "depth": d}, "test")
    assert y.shape == (B, T, C) and f.last_indices[0].shape == (B, T // 4)
    np.testing.assert_allclose(f.last_erank[0].cpu().numpy(), EO.erank(rgb.numpy()), rtol=1e-4)
    y.sum().backward()
    assert torch.isfinite(r.grad).all() and torch.isfinite(d.grad).all()
    with pytest.raises(ValueError):
        r3d_b200.CMFuser(C, variant="vary", select_axis="token")


# ------------------------------------------------------------------ f1: tcgen05 linear GEMM with fused epilogues
def _gelu_ref(x):
    return torch.nn.functional.gelu(x)


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("M,N,K", [(256, 512, 512), (300, 136, 72), (1024, 2048, 512), (128, 64, 64)])
def test_gemm_forward_epilogues(M, N, K, dt, dev):
    from r3d_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    X = (torch.randn(M, K, generator=g) / K ** 0.5).to(dt).to(dev)
    W = torch.randn(N, K, generator=g).to(dt).to(dev)
    b = torch.randn(N, generator=g).to(dt).to(dev)
    R = torch.randn(M, N, generator=g).to(dt).to(dev)
    tol = dict(rtol=2e-2, atol=2e-2) if dt == torch.bfloat16 else dict(rtol=2e-5, atol=5e-5)
    ref = X.double() @ W.double().T
    D = ops.gemm(X, W)
    torch.testing.assert_close(D.double(), ref, **tol)
    # bias + GELU with the pre-activation saved, residual, column sums of the stored result
    D, H, cs = ops.gemm(X, W, bias=b, act=ops.ACT_GELU, residual=R, want_aux=True, colsum=True)
    pre = ref + b.double()
    torch.testing.assert_close(H.double(), pre, **tol)
    want = _gelu_ref(H.double()) + R.double()              # GELU of the STORED pre-activation (what autograd would see)
    torch.testing.assert_close(D.double(), want, **tol)
    torch.testing.assert_close(cs.double(), D.double().sum(0), rtol=1e-4, atol=1e-3 * M ** 0.5)
    # ReLU + |.| column sums (the channel-score epilogue of the input projections)
    D, cs = ops.gemm(X, W, bias=b, act=ops.ACT_RELU, colsum=True, colsum_abs=True)
    torch.testing.assert_close(D.double(), torch.relu(pre), **tol)
    torch.testing.assert_close(cs.double(), D.double().abs().sum(0), rtol=1e-4, atol=1e-3 * M ** 0.5)


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("M,N,K", [(512, 256, 128), (4096, 512, 512), (264, 72, 136)])
def test_gemm_backward_forms(M, N, K, dt, dev):
    """dX = dY W (B operand MN-major) with the gelu' epilogue, dW = dY^T X (both MN-major, split along K)."""
    from r3d_b200 import ops
    g = torch.Generator().manual_seed(7 * M + N)
    X = torch.randn(M, K, generator=g).to(dt).to(dev)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(dt).to(dev)
    dY = (torch.randn(M, N, generator=g) / N ** 0.5).to(dt).to(dev)
    H = torch.randn(M, K, generator=g).to(dt).to(dev)
    tol = dict(rtol=2e-2, atol=2e-2) if dt == torch.bfloat16 else dict(rtol=2e-5, atol=5e-5)
    dX = ops.gemm(dY, W, True, False)
    torch.testing.assert_close(dX.double(), dY.double() @ W.double(), **tol)
    dXg, cs = ops.gemm(dY, W, True, False, aux_in=H, colsum=True)
    hd = H.double().requires_grad_(True)
    _gelu_ref(hd).sum().backward()
    torch.testing.assert_close(dXg.double(), (dY.double() @ W.double()) * hd.grad, **tol)
    torch.testing.assert_close(cs.double(), dXg.double().sum(0), rtol=1e-4, atol=1e-3 * M ** 0.5)
    dW = ops.gemm(dY, X, False, False)
    ref = dY.double().T @ X.double()
    tolw = dict(rtol=2e-2, atol=2e-2 * M ** 0.5 / 8) if dt == torch.bfloat16 else dict(rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(dW.double(), ref, **tolw)
    torch.testing.assert_close(ops.colsum(dY).double(), dY.double().sum(0), rtol=2e-2 if dt == torch.bfloat16 else 1e-5,
                               atol=1e-2 if dt == torch.bfloat16 else 1e-4)


def test_fuser_train_step_vs_oracle(dev):
    """ops.FuserTrainStep (what bench.py times): fused output and input gradients = CMFuser backward + d(mean erank)/dX,
    against the torch port (autograd) + the float64 erank oracle; parameter gradients against the port."""
    import r3d_b200
    from r3d_b200 import ops
    from oracle.torch_port import PortCMFuser
    B, T, C = 3, 32, 64
    rgb, dep = synth(B, T, C, 21)
    gy = torch.randn(B, T, C, generator=torch.Generator().manual_seed(8))
    torch.manual_seed(0)
    ref = PortCMFuser(C, depth=1, num_heads=4, variant="tokenfusion").train()
    ref.embd_drop.p = 0.0
    f = r3d_b200.CMFuser(C, depth=1, num_heads=4, score_scope="global")
    f.load_state_dict(ref.state_dict())
    f = f.to(dev).train()
    f.embd_drop.p = 0.0
    step = ops.FuserTrainStep(f, B, T, C, torch.float32, dev)
    buf = torch.stack([rgb, dep]).to(dev)
    y, er, gr, gd = step(buf, gy.to(dev))
    r = rgb.clone().requires_grad_(True)
    d = dep.clone().requires_grad_(True)
    yr = ref({"rgb": r, "depth": d}, "test")
    yr.backward(gy)
    w = np.full(B, 1.0 / (2 * B))
    g_r = r.grad.numpy() + EO.erank_bwd(rgb.numpy(), w)
    g_d = d.grad.numpy() + EO.erank_bwd(dep.numpy(), w)
    np.testing.assert_allclose(y.detach().cpu().numpy(), yr.detach().numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(er.cpu().numpy(), np.concatenate([EO.erank(rgb.numpy()), EO.erank(dep.numpy())]), rtol=1e-4)
    np.testing.assert_allclose(gr.cpu().numpy(), g_r, rtol=1e-3, atol=2e-5)
    np.testing.assert_allclose(gd.cpu().numpy(), g_d, rtol=1e-3, atol=2e-5)
    for (n, p), (_, q) in zip(f.named_parameters(), ref.named_parameters()):
        if q.grad is not None and p.grad is not None:
            np.testing.assert_allclose(p.grad.cpu().numpy(), q.grad.numpy(), rtol=1e-3,
                                       atol=1e-4 * max(1.0, q.grad.abs().max().item()), err_msg=n)
    # the packed statistic carries the erank sum of the step
    np.testing.assert_allclose(f.last_packed[2 * C].item(), er.sum().item(), rtol=1e-5)


# ------------------------------------------------------------------ f2 / f3: input projections with the score by-product
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_rgb_and_depth_embed_vs_torch(dt, dev):
    """relu(Linear) (tokenfusion.py:179-183) and relu(LN(Linear)) (:194-197) against torch in float64, forward, all
    gradients, and the |output| column sums that replace the fuser's score pass."""
    import r3d_b200
    B, S, K, HW, C = 3, 40, 256, 24 * 24, 128
    g = torch.Generator().manual_seed(4)
    feats = torch.randn(B, S, K, generator=g)
    depth = torch.randn(B, S, 24, 24, generator=g)
    torch.manual_seed(1)
    rgbm, depm = r3d_b200.RGBEmbed(K, C), r3d_b200.DepthEmbed(HW, C)
    with torch.no_grad():
        depm.depth_layernorm.weight.uniform_(0.5, 1.5)
        depm.depth_layernorm.bias.uniform_(-0.3, 0.3)
    ref_r, ref_d = [torch.nn.Linear(K, C).double(), torch.nn.Linear(HW, C).double()]
    ref_ln = torch.nn.LayerNorm(C).double()
    ref_r.load_state_dict({k: v.double() for k, v in rgbm.input_embed.state_dict().items()})
    ref_d.load_state_dict({k: v.double() for k, v in depm.depth_projection.state_dict().items()})
    ref_ln.load_state_dict({k: v.double() for k, v in depm.depth_layernorm.state_dict().items()})
    rgbm, depm = rgbm.to(dev).to(dt), depm.to(dev).to(dt)
    if dt == torch.bfloat16:          # the reference computes on the rounded parameters / inputs too
        for m_, r_ in ((rgbm.input_embed, ref_r), (depm.depth_projection, ref_d), (depm.depth_layernorm, ref_ln)):
            r_.load_state_dict({k: v.double().cpu() for k, v in m_.state_dict().items()})
    f = feats.to(dt).to(dev).requires_grad_(True)
    d = depth.to(dt).to(dev).requires_grad_(True)
    y_r, y_d = rgbm(f), depm(d)
    gy = torch.randn(B, S, C, generator=g)
    (y_r * gy.to(dt).to(dev)).sum().backward()
    (y_d * gy.to(dt).to(dev)).sum().backward()
    f64 = feats.to(dt).double().requires_grad_(True)
    d64 = depth.to(dt).double().requires_grad_(True)
    yr = torch.relu(ref_r(f64))
    yd = torch.relu(ref_ln(ref_d(d64.view(B, S, -1))))
    (yr * gy.to(dt).double()).sum().backward()
    (yd * gy.to(dt).double()).sum().backward()
    tol = dict(rtol=3e-2, atol=3e-2) if dt == torch.bfloat16 else dict(rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(y_r.double().cpu(), yr.detach(), **tol)
    torch.testing.assert_close(y_d.double().cpu(), yd.detach(), **tol)
    if dt == torch.bfloat16:
        # a pre-activation that rounds across zero in bf16 flips its ReLU mask relative to the float64 reference: a few
        # rows differ by O(1) terms, so the input gradients are compared norm-wise
        for got, ref in ((f.grad.double().cpu(), f64.grad), (d.grad.double().cpu().view(B, S, -1), d64.grad.view(B, S, -1))):
            assert (got - ref).norm() <= 5e-2 * ref.norm()
    else:
        torch.testing.assert_close(f.grad.double().cpu(), f64.grad, **tol)
        torch.testing.assert_close(d.grad.double().cpu().view(B, S, -1), d64.grad.view(B, S, -1), **tol)
    for ours, ref in ((rgbm.input_embed, ref_r), (depm.depth_projection, ref_d), (depm.depth_layernorm, ref_ln)):
        for (n, p), (_, q) in zip(ours.named_parameters(), ref.named_parameters()):
            if dt == torch.bfloat16:
                assert (p.grad.double().cpu() - q.grad).norm() <= 5e-2 * q.grad.norm() + 1e-2, n
            else:
                torch.testing.assert_close(p.grad.double().cpu(), q.grad, rtol=1e-4, atol=1e-3, msg=lambda m, n=n: f"{n}: {m}")
    # score by-products = column sums of |stored output|
    torch.testing.assert_close(rgbm.last_score.sums().double().cpu(), y_r.detach().double().abs().sum((0, 1)).cpu(),
                               rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(depm.last_score.sums().double().cpu(), y_d.detach().double().abs().sum((0, 1)).cpu(),
                               rtol=1e-4, atol=1e-3)


def test_fuser_front_skips_score_pass(dev):
    """FuserFront (tokenfusion.py:179-199 wiring): the channel selection fed by the producers' by-products equals the
    selection of the stand-alone score pass, and the fused output equals CMFuser on the embedded tensors."""
    import r3d_b200
    from r3d_b200 import _lib
    B, S, K, HW, C = 2, 48, 128, 16 * 16, 64
    g = torch.Generator().manual_seed(9)
    feats, depth = torch.randn(B, S, K, generator=g).to(dev), torch.randn(B, S, 16, 16, generator=g).to(dev)
    torch.manual_seed(3)
    front = r3d_b200.FuserFront(K, HW, C, n_head=4).to(dev).eval()
    y = front(feats, depth, "test")
    idx = [i.clone() for i in front.fuser.last_indices]
    src, dep = front.rgb(feats), front.depth(depth)
    y2 = front.fuser({"rgb": src, "depth": dep}, "test")               # own score pass
    assert torch.equal(idx[0], front.fuser.last_indices[0]) and torch.equal(idx[1], front.fuser.last_indices[1])
    torch.testing.assert_close(y, y2, rtol=1e-5, atol=1e-6)
    ref_idx = O.bottomk(O.channel_score(src.detach().cpu().numpy()), C // 4)
    np.testing.assert_array_equal(idx[0].cpu().numpy(), ref_idx)


# ------------------------------------------------------------------ N2: M-modality fuser (rgb + depth + gaze)
@pytest.mark.parametrize("M", [3, 4])
def test_multi_modality_fuser_vs_oracle(M, dev):
    """CMFuser with M >= 3 modal features (configs[4]) against oracle/torch_port.py:PortCMFuserM (the reference's
    modules with the M x M -inf-diagonal mask; parity unpinned beyond M = 2): selected channels bit-exact, output and all
    gradients within the fp32 bars."""
    import r3d_b200
    from oracle.torch_port import PortCMFuserM
    B, T, C, H = 2, 24, 128, 4                      # head_dim 32
    g = torch.Generator().manual_seed(40 + M)
    c = torch.arange(C, dtype=torch.float32)
    names = ["rgb", "depth", "gaze", "audio"][:M]
    feats = {}
    for i, n in enumerate(names):
        perm = torch.randperm(C, generator=g)
        feats[n] = (torch.relu(torch.randn(B, T, C, generator=g)) * (1 + (i + 1) * c / C))[:, :, perm].contiguous()
    gy = torch.randn(B, T, C, generator=g)
    torch.manual_seed(0)
    ref = PortCMFuserM(C, depth=1, num_heads=H).train()
    ref.embd_drop.p = 0.0
    f = r3d_b200.CMFuser(C, depth=1, num_heads=H)
    f.load_state_dict(ref.state_dict())
    f = f.to(dev).train()
    f.embd_drop.p = 0.0
    ours_in = {n: v.to(dev).requires_grad_(True) for n, v in feats.items()}
    ref_in = {n: v.clone().requires_grad_(True) for n, v in feats.items()}
    y = f(ours_in, "test")
    yr = ref(ref_in, "test")
    for m in range(M):
        np.testing.assert_array_equal(f.last_indices[m].cpu().numpy(), ref.last_indices[m].numpy())
    np.testing.assert_allclose(y.detach().cpu().numpy(), yr.detach().numpy(), rtol=1e-4, atol=1e-5)
    y.backward(gy.to(dev))
    yr.backward(gy)
    for n in names:
        np.testing.assert_allclose(ours_in[n].grad.cpu().numpy(), ref_in[n].grad.numpy(), rtol=1e-3, atol=2e-5, err_msg=n)
    for (n, p), (_, q) in zip(f.named_parameters(), ref.named_parameters()):
        if q.grad is not None:
            assert p.grad is not None, n
            np.testing.assert_allclose(p.grad.cpu().numpy(), q.grad.numpy(), rtol=1e-3,
                                       atol=1e-4 * max(1.0, q.grad.abs().max().item()), err_msg=n)


def test_multi_modality_kernels_bf16_and_m2_equivalence(dev):
    """exchange_multi with M = 2 equals the reference swap; the M-token attention in bf16 against torch's masked
    softmax attention in float64; token_mean against torch."""
    from r3d_b200 import ops
    B, T, C, H, M = 2, 16, 256, 8, 3
    rgb, dep = synth(B, T, C, 50)
    idx = torch.stack([torch.randperm(C)[:C // 4], torch.randperm(C)[:C // 4]]).to(dev)
    a = ops.exchange_multi([rgb.to(dev), dep.to(dev)], idx)
    b = ops.exchange(rgb.to(dev), dep.to(dev), idx[0], idx[1])
    assert torch.equal(a, b)
    g = torch.Generator().manual_seed(3)
    qkv = torch.randn(B * T, M, 3 * C, generator=g).to(torch.bfloat16)
    q = qkv.to(dev).requires_grad_(True)
    out = ops.mtoken_attention(q, H)
    go = torch.randn(B * T, M, C, generator=g).to(torch.bfloat16)
    out.backward(go.to(dev))
    q64 = qkv.double().requires_grad_(True)
    hd = C // H
    t = q64.view(B * T, M, 3, H, hd).permute(2, 0, 3, 1, 4)
    w = (t[0] @ t[1].transpose(-2, -1)) * hd ** -0.5 + torch.zeros(M, M).masked_fill(torch.eye(M) == 1, float("-inf"))
    ref = (w.softmax(-1) @ t[2]).transpose(1, 2).reshape(B * T, M, C)
    ref.backward(go.double())
    torch.testing.assert_close(out.double().cpu(), ref.detach(), rtol=2e-2, atol=2e-2)
    assert (q.grad.double().cpu() - q64.grad).norm() <= 2e-2 * q64.grad.norm()
    x = torch.randn(40, M, C, generator=g).to(dev).requires_grad_(True)
    m = ops.token_mean(x)
    torch.testing.assert_close(m, x.mean(1))
    m.sum().backward()
    torch.testing.assert_close(x.grad, torch.full_like(x, 1.0 / M))


# ------------------------------------------------------------------ f4: FUTR around the fuser path
def test_futr_matches_reference_module(dev):
    """r3d_b200.FUTR against the UNMODIFIED reference FUTR (oracle/_ref or /root/reference, run on the CPU) with the same
    weights and inputs: identical state_dict names (strict load both ways), eval outputs within fp32 tolerance."""
    import types
    import r3d_b200
    from oracle import ref_loader
    RefFUTR = ref_loader.load_futr("tokenfusion")
    if RefFUTR is None:
        pytest.skip("no reference tree (oracle/_ref not staged)")
    args = types.SimpleNamespace(input_dim=256, seg=True, anticipate=True, max_pos_len=64, input_type="i3d_transcript")
    torch.manual_seed(0)
    ref = RefFUTR(n_class=12, hidden_dim=64, src_pad_idx=13, device="cpu", args=args, n_query=8, n_head=4,
                  num_encoder_layers=1, num_decoder_layers=2, query_num=20).eval()
    ours = r3d_b200.FUTR(n_class=12, hidden_dim=64, src_pad_idx=13, device=dev, args=args, n_query=8, n_head=4,
                         num_encoder_layers=1, num_decoder_layers=2, query_num=20)
    assert set(ours.state_dict().keys()) == set(ref.state_dict().keys())
    ours.load_state_dict(ref.state_dict(), strict=True)
    ours = ours.to(dev).eval()
    g = torch.Generator().manual_seed(2)
    B, S = 2, 12
    feats = torch.randn(B, S, 256, generator=g)
    depth = torch.randn(B, S, 224, 224, generator=g) * 0.1
    with torch.no_grad():
        want = ref((feats, None), depth, mode="test")
        got = ours((feats.to(dev), None), depth.to(dev), mode="test")
    assert set(got) == set(want) == {"action", "duration", "seg"}
    for k in want:
        np.testing.assert_allclose(got[k].cpu().numpy(), want[k].numpy(), rtol=2e-3, atol=2e-4, err_msg=k)
    # and it trains: gradients reach the input projections through the fuser
    ours.train()
    out = ours((feats.to(dev), torch.zeros(B, S, dtype=torch.long, device=dev)), depth.to(dev), mode="train")
    (out["action"].sum() + out["duration"].sum() + out["seg"].sum()).backward()
    assert ours.input_embed.weight.grad is not None and ours.depth_projection.weight.grad is not None
    assert torch.isfinite(ours.depth_projection.weight.grad).all()


def test_torch_ops_schemas_with_autograd(dev):
    """torch.ops.r3d.* (SURVEY.md 8b): exchange_fwd / exchange_bwd / erank_fwd / erank_bwd are registered with autograd
    and agree with the Python-level operators."""
    from r3d_b200 import ops
    B, T, C = 2, 16, 64
    rgb, dep = synth(B, T, C, 12)
    r = rgb.to(dev).requires_grad_(True)
    d = dep.to(dev).requires_grad_(True)
    idx = ops.bottomk(ops.channel_score(r.detach(), d.detach()), C // 4)
    alpha = torch.rand(1, 1, C, device=dev, requires_grad=True)
    out = torch.ops.r3d.exchange_fwd(r, d, idx[0], idx[1], alpha, ops.BLEND_SCALE)
    g = torch.randn_like(out)
    out.backward(g)
    r2, d2, a2 = rgb.to(dev).requires_grad_(True), dep.to(dev).requires_grad_(True), alpha.detach().clone().requires_grad_(True)
    ops.exchange(r2, d2, idx[0], idx[1], a2, ops.BLEND_SCALE).backward(g)
    assert torch.equal(r.grad, r2.grad) and torch.equal(d.grad, d2.grad)
    torch.testing.assert_close(alpha.grad, a2.grad)
    x = torch.from_numpy(_spectra("relu", 2, 64, 128, 5)).to(dev).requires_grad_(True)
    er, sigma, U, Y = torch.ops.r3d.erank_fwd(x, 1e-4)
    er.sum().backward()
    x2 = x.detach().clone().requires_grad_(True)
    ops.erank(x2).sum().backward()
    assert torch.equal(x.grad, x2.grad)


def test_fuser_step_host_entry(dev):
    """r3d_fuser_step_host: the whole hot path from host buffers through ONE C entry, against the oracle."""
    from r3d_b200 import ops
    B, T, C = 3, 40, 64
    rgb, dep = synth(B, T, C, 77)
    gst = torch.randn(B, T, 2, C, generator=torch.Generator().manual_seed(2))
    st, er, d_r, d_d, idx = ops.fuser_step_host(rgb, dep, gst)
    ref, ir, idd = O.token_fusion("tokenfusion", rgb.numpy(), dep.numpy(), "test", return_indices=True)
    np.testing.assert_array_equal(idx[0].numpy(), ir)
    np.testing.assert_array_equal(idx[1].numpy(), idd)
    np.testing.assert_array_equal(st.numpy(), ref)
    np.testing.assert_allclose(er.numpy(), np.concatenate([EO.erank(rgb.numpy()), EO.erank(dep.numpy())]), rtol=1e-4)
    gr, gd, _ = O.exchange_bwd(gst.numpy(), rgb.numpy(), dep.numpy(), ir, idd)
    w = np.full(B, 1.0 / (2 * B))
    np.testing.assert_allclose(d_r.numpy(), gr + EO.erank_bwd(rgb.numpy(), w), rtol=1e-3, atol=2e-5)
    np.testing.assert_allclose(d_d.numpy(), gd + EO.erank_bwd(dep.numpy(), w), rtol=1e-3, atol=2e-5)
    st2, er2, n1, n2, _ = ops.fuser_step_host(rgb, dep)           # forward only
    assert n1 is None and n2 is None and torch.equal(st2, st)
