"""CPU oracle for effective rank -- TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.

The reference repository contains no effective-rank code (no svd,
eigh, Gram or spectral entropy anywhere under /root/reference; README.md:13
describes it in prose only).  This file restates the standard definition
(Roy & Vetterli 2007) in float64 numpy so the CUDA chain
(Gram -> Jacobi -> Rayleigh refinement -> entropy/exp) has something to be
checked against.  There are no reference vectors to pin it to; DESIGN.md says
the same.

Conventions this build fixes (the reference pins none of them):
  * no centering; natural log; 0 * ln 0 := 0;
  * n = min(T, C) singular values of the (T, C) sample;
  * numerical-rank cut-off: singular values with sigma_j <= rtol * sigma_max are
    treated as exactly zero (p_j = 0).  Default rtol = 1e-4 -- the level an fp32
    Gram can resolve ("erank accuracy");
  * batch statistic = arithmetic mean of the per-sample values.
"""
from __future__ import annotations

import numpy as np

DEFAULT_RTOL = 1e-4


def singular_values(x: np.ndarray) -> np.ndarray:
    """(B, T, C) -> (B, n) float64, descending (numpy.linalg.svd order)."""
    return np.linalg.svd(np.asarray(x, dtype=np.float64), compute_uv=False)


def erank_from_sigma(sigma: np.ndarray, rtol: float = DEFAULT_RTOL):
    """sigma (B, n) -> (erank (B,), H (B,), S (B,), keep (B, n) bool)."""
    sigma = np.asarray(sigma, dtype=np.float64)
    smax = sigma.max(axis=-1, keepdims=True)
    keep = sigma > rtol * smax
    s = np.where(keep, sigma, 0.0)
    S = s.sum(axis=-1, keepdims=True)
    S_safe = np.where(S > 0, S, 1.0)
    p = s / S_safe
    with np.errstate(divide="ignore", invalid="ignore"):
        plogp = np.where(p > 0, p * np.log(np.where(p > 0, p, 1.0)), 0.0)
    H = -plogp.sum(axis=-1)
    er = np.exp(H)
    er = np.where(S[..., 0] > 0, er, 0.0)   # all-zero sample: erank := 0
    return er, H, S[..., 0], keep


def erank(x: np.ndarray, rtol: float = DEFAULT_RTOL) -> np.ndarray:
    """Per-sample effective rank of x (B, T, C): exp(-sum p ln p), p = sigma / sum sigma."""
    return erank_from_sigma(singular_values(x), rtol)[0]


def erank_gram_route(x: np.ndarray, rtol: float = DEFAULT_RTOL) -> np.ndarray:
    """Same quantity through the route the CUDA path takes: Gram on the smaller
    side -> symmetric eigendecomposition -> sigma = sqrt(max(lambda, 0)).  Used by
    the tests to show the two routes agree in float64."""
    x = np.asarray(x, dtype=np.float64)
    B, T, C = x.shape
    G = x @ x.transpose(0, 2, 1) if T <= C else x.transpose(0, 2, 1) @ x
    lam = np.linalg.eigvalsh(G)[:, ::-1]
    return erank_from_sigma(np.sqrt(np.clip(lam, 0.0, None)), rtol)[0]


def erank_bwd(x: np.ndarray, g: np.ndarray, rtol: float = DEFAULT_RTOL) -> np.ndarray:
    """d(sum_b g_b * erank_b) / dx, float64.

    With S = sum sigma, p = sigma / S, H = -sum p ln p:
        d erank / d sigma_j = erank * (-(ln p_j + H) / S)
        d sigma_j / d X     = u_j v_j^T
    Cut-off singular values (p_j = 0) carry no gradient."""
    x = np.asarray(x, dtype=np.float64)
    g = np.asarray(g, dtype=np.float64)
    U, s, Vt = np.linalg.svd(x, full_matrices=False)
    er, H, S, keep = erank_from_sigma(s, rtol)
    S_safe = np.where(S > 0, S, 1.0)[:, None]
    p = np.where(keep, s, 0.0) / S_safe
    with np.errstate(divide="ignore", invalid="ignore"):
        lnp = np.where(p > 0, np.log(np.where(p > 0, p, 1.0)), 0.0)
    dsig = np.where(keep, er[:, None] * (-(lnp + H[:, None]) / S_safe), 0.0)
    coef = g[:, None] * dsig
    return np.einsum("bij,bj,bjk->bik", U, coef, Vt)


def token_informativeness(x: np.ndarray, rtol: float = DEFAULT_RTOL) -> np.ndarray:
    """Per-token informativeness derived from the spectrum (SURVEY.md a13 -- no
    reference symbol, unpinned): the leverage of token t weighted by the
    normalised spectrum,  s_t = sum_j p_j * u_{tj}^2  (rows of U for T <= C,
    rows of V otherwise index the *other* axis, so the score is always over the
    n = min(T, C) side).  Sums to 1 over that side."""
    x = np.asarray(x, dtype=np.float64)
    B, T, C = x.shape
    U, s, Vt = np.linalg.svd(x, full_matrices=False)
    er, H, S, keep = erank_from_sigma(s, rtol)
    p = np.where(keep, s, 0.0) / np.where(S > 0, S, 1.0)[:, None]
    if T <= C:
        return np.einsum("btj,bj->bt", U ** 2, p)
    return np.einsum("bjc,bj->bc", Vt ** 2, p)


def token_scores(x: np.ndarray, rtol: float = DEFAULT_RTOL) -> np.ndarray:
    """(B, T): informativeness of every TOKEN, whichever side is shorter -- s_t = sum_j p_j u_{tj}^2 with u_j the left
    singular vectors of the (T, C) sample (north_star kernel 3; no reference symbol, unpinned).  Sums to 1 over t."""
    x = np.asarray(x, dtype=np.float64)
    U, s, Vt = np.linalg.svd(x, full_matrices=False)
    er, H, S, keep = erank_from_sigma(s, rtol)
    p = np.where(keep, s, 0.0) / np.where(S > 0, S, 1.0)[:, None]
    return np.einsum("btj,bj->bt", U ** 2, p)
