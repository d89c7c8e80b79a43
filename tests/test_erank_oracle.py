"""Pin and cross-check the effective-rank oracle on the CPU.

The reference has no effective-rank code, so `oracle/erank_oracle.py` cannot be pinned to reference
outputs.  What CAN be done, and is done here:
  * fixtures `tests/golden/erank_*.npz` come from an independently written torch float64 implementation
    (`tests/golden/make_erank_golden.py`: `torch.linalg.svdvals` + autograd); the numpy oracle's forward AND its
    hand-derived gradient must reproduce them;
  * the two float64 routes of the oracle (svd of X, eigenvalues of the Gram) must agree;
  * `torch.autograd.gradcheck` on the differentiable torch restatement;
  * `hypothesis` property tests of the definition: range, scale / orthogonal / permutation
    invariance, known spectra;
  * the behaviour of the numerical-rank cut-off ACROSS its threshold (the one convention appendix B does not have).
"""
import glob
import hashlib
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import erank_oracle as EO
from oracle.torch_port import erank_torch

sys.path.insert(0, GOLDEN)
from make_erank_golden import make_input  # noqa: E402

FIX = sorted(glob.glob(os.path.join(GOLDEN, "erank_*.npz")))


def _load(path):
    z = np.load(path)
    x = make_input(str(z["kind"]), int(z["B"]), int(z["T"]), int(z["C"]), int(z["seed"]))
    assert hashlib.sha256(x.tobytes()).hexdigest() == str(z["x_sha256"]), "numpy RNG stream drifted: regenerate fixtures"
    return z, x


def test_fixtures_present():
    assert len(FIX) == 12, FIX


@pytest.mark.parametrize("path", FIX, ids=os.path.basename)
def test_oracle_forward_matches_torch_float64(path):
    z, x = _load(path)
    np.testing.assert_allclose(EO.singular_values(x), z["sigma"], rtol=1e-9, atol=1e-9 * z["sigma"].max())
    np.testing.assert_allclose(EO.erank(x), z["erank"], rtol=1e-10)
    np.testing.assert_allclose(EO.erank(x, rtol=0.0), z["erank_rtol0"], rtol=1e-10)


@pytest.mark.parametrize("path", FIX, ids=os.path.basename)
def test_oracle_gram_route_agrees(path):
    """svd(X) and sqrt(eigvalsh(Gram)) in float64.  The Gram squares the condition number: singular values below
    ~1e-8 sigma_max come out of the Gram route as noise, so the exactly rank-deficient fixtures are compared with the
    cut-off (which removes them) and the full-rank ones also without."""
    z, x = _load(path)
    np.testing.assert_allclose(EO.erank_gram_route(x), EO.erank(x), rtol=1e-7)
    if str(z["kind"]) != "rankdef":
        np.testing.assert_allclose(EO.erank_gram_route(x, rtol=0.0), EO.erank(x, rtol=0.0), rtol=1e-6)


@pytest.mark.parametrize("path", FIX, ids=os.path.basename)
def test_oracle_gradient_matches_autograd(path):
    """The hand-derived gradient (oracle/erank_oracle.py:erank_bwd) against float64 autograd
    through torch.linalg.svdvals (stored by make_erank_golden.py)."""
    z, x = _load(path)
    g = EO.erank_bwd(x, np.ones(int(z["B"])))
    ref = z["grad"].astype(np.float64)
    assert np.abs(g - ref).max() <= 2e-6 * np.abs(ref).max()      # fixture is stored in float32


def test_oracle_gradient_live_autograd_and_weights():
    """Same check live (no fixture), with non-uniform upstream weights and both sides (T < C, T > C)."""
    rng = np.random.default_rng(5)
    for (B, T, C) in ((3, 12, 20), (2, 24, 10), (2, 16, 16)):
        x = np.maximum(rng.standard_normal((B, T, C)), 0).astype(np.float32)
        w = rng.standard_normal(B)
        xt = torch.from_numpy(x).double().requires_grad_(True)
        (erank_torch(xt) * torch.from_numpy(w)).sum().backward()
        np.testing.assert_allclose(EO.erank_bwd(x, w), xt.grad.numpy(), rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(EO.erank(x), erank_torch(xt).detach().numpy(), rtol=1e-12)


def test_torch_restatement_gradcheck():
    torch.manual_seed(0)
    x = torch.randn(2, 6, 9, dtype=torch.float64, requires_grad=True)
    assert torch.autograd.gradcheck(lambda t: erank_torch(t), (x,), eps=1e-6, atol=1e-6, rtol=1e-5)
    y = torch.randn(2, 9, 5, dtype=torch.float64, requires_grad=True)
    assert torch.autograd.gradcheck(lambda t: erank_torch(t), (y,), eps=1e-6, atol=1e-6, rtol=1e-5)


# ---------------------------------------------------------------------------------------------------------
# properties of the definition
# ---------------------------------------------------------------------------------------------------------
hyp = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st  # noqa: E402
from hypothesis.extra import numpy as hnp  # noqa: E402

_shape = st.tuples(st.integers(1, 3), st.integers(1, 12), st.integers(1, 12))
_mats = _shape.flatmap(lambda s: hnp.arrays(np.float64, s, elements=st.floats(-4, 4, allow_nan=False, width=32)))


@settings(max_examples=60, deadline=None)
@given(_mats)
def test_property_range_and_scale_invariance(x):
    er = EO.erank(x, rtol=0.0)
    n = min(x.shape[1:])
    nz = np.abs(x).reshape(x.shape[0], -1).max(axis=1) > 0
    assert np.all(er[nz] >= 1 - 1e-9) and np.all(er[nz] <= n + 1e-9)
    assert np.all(er[~nz] == 0)                                   # all-zero sample: erank := 0
    np.testing.assert_allclose(EO.erank(3.7 * x, rtol=0.0), er, rtol=1e-9)
    np.testing.assert_allclose(EO.erank(x.transpose(0, 2, 1), rtol=0.0), er, rtol=1e-9)   # same singular values


@settings(max_examples=40, deadline=None)
@given(_mats, st.integers(0, 2 ** 31 - 1))
def test_property_orthogonal_and_permutation_invariance(x, seed):
    rng = np.random.default_rng(seed)
    B, T, C = x.shape
    qt, _ = np.linalg.qr(rng.standard_normal((T, T)))
    qc, _ = np.linalg.qr(rng.standard_normal((C, C)))
    er = EO.erank(x)
    np.testing.assert_allclose(EO.erank(qt @ x @ qc), er, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(EO.erank(x[:, rng.permutation(T)][:, :, rng.permutation(C)]), er, rtol=1e-9, atol=1e-12)


def test_known_spectra():
    rng = np.random.default_rng(0)
    u = rng.standard_normal((1, 9, 1)); v = rng.standard_normal((1, 1, 14))
    np.testing.assert_allclose(EO.erank(u @ v), [1.0], rtol=1e-9)                      # rank one
    q, _ = np.linalg.qr(rng.standard_normal((14, 9)))
    np.testing.assert_allclose(EO.erank(q.T[None]), [9.0], rtol=1e-9)                  # orthonormal rows: erank = n
    s = np.array([4.0, 2.0, 1.0, 1.0])
    p = s / s.sum()
    np.testing.assert_allclose(EO.erank(np.diag(s)[None]), [np.exp(-(p * np.log(p)).sum())], rtol=1e-12)


def test_cutoff_across_threshold():
    """The numerical-rank cut-off sigma_j <= rtol * sigma_max -> 0 (default 1e-4) is this build's convention, not
    appendix B's ("clamp inside the log only").  It makes erank discontinuous where a singular value crosses the
    threshold; this test pins HOW discontinuous: dropping one singular value sigma = t * sigma_max changes the entropy
    by at most  p |ln p| + p  with  p = t sigma_max / S <= t,  i.e. <= 1.1e-3 relative in erank at t = 1e-4 in the worst
    case S = sigma_max, and ~1e-5 for the spectra of the metric (S / sigma_max ~ 20..60).  Above the threshold the two
    conventions are identical; on exactly rank-deficient inputs they agree in float64 (null singular values ~1e-16)."""
    base = np.array([1.0, 0.8, 0.5, 0.3, 0.1])
    for t in (0.5e-4, 0.99e-4, 1.01e-4, 2e-4):
        s = np.concatenate([base, [t]])
        x = np.diag(s)[None]
        cut, none = EO.erank(x, rtol=1e-4)[0], EO.erank(x, rtol=0.0)[0]
        p = t / s.sum()
        bound = (p * abs(np.log(p)) + p) * 1.01
        if t > 1e-4:
            assert cut == none                                     # kept: conventions coincide
        else:
            assert 0 < (none - cut) / none <= bound                # dropped: bounded jump
    # the metric's own spectra (fixtures): the two conventions differ by < 3e-4 relative, and only where the
    # spectrum reaches below the cut-off (channel-decay 512 x 512)
    for path in FIX:
        z = np.load(path)
        rel = np.abs(z["erank"] - z["erank_rtol0"]) / z["erank_rtol0"]
        assert rel.max() < 3e-4
        if str(z["kind"]) in ("relu", "gauss") and int(z["T"]) < int(z["C"]):
            assert rel.max() < 1e-13


def test_token_axis_oracle_against_torch_autograd():
    """oracle/fuser_oracle.py:token_fusion_tokens / token_exchange_bwd (unpinned restatement) against the same
    operation written with torch index assignment + autograd; scores sum to one and pick the weakest tokens."""
    from oracle import fuser_oracle as O
    rng = np.random.default_rng(3)
    B, T, C = 3, 12, 7
    rgb = np.maximum(rng.standard_normal((B, T, C)), 0).astype(np.float32)
    dep = np.maximum(rng.standard_normal((B, T, C)), 0).astype(np.float32)
    rgb[:, 5] *= 1e-3                      # an uninformative token must be selected
    s = EO.token_scores(rgb)
    np.testing.assert_allclose(s.sum(1), 1.0, rtol=1e-12)
    out, ir, idd = O.token_fusion_tokens(rgb, dep, return_indices=True)
    assert all(5 in ir[b] for b in range(B)) and ir.shape == (B, T // 4)
    r = torch.from_numpy(rgb).requires_grad_(True)
    d = torch.from_numpy(dep).requires_grad_(True)
    ex_r, ex_d = r.clone(), d.clone()
    for b in range(B):
        ex_r[b, torch.from_numpy(ir[b])] = d[b, torch.from_numpy(ir[b])]
        ex_d[b, torch.from_numpy(idd[b])] = r[b, torch.from_numpy(idd[b])]
    st = torch.stack([ex_r, ex_d], dim=2)
    np.testing.assert_array_equal(st.detach().numpy(), out)
    g = torch.from_numpy(rng.standard_normal((B, T, 2, C)).astype(np.float32))
    st.backward(g)
    gr, gd = O.token_exchange_bwd(g.numpy(), ir, idd)
    np.testing.assert_array_equal(r.grad.numpy(), gr)
    np.testing.assert_array_equal(d.grad.numpy(), gd)
