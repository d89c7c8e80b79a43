"""CPU oracle for the R3D token-fuser hot path -- TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of what the reference computes on the path
SURVEY.md section 8 names.  It is the *checker*: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it.  Nothing under ``r3d_b200/`` imports it and the
product path has no CPU fallback.

Parity status:
  * channel score, bottom-k, exchange / alpha-scaled exchange / convex blend,
    BatchNorm front end, their backward, and the CMFuser wrapper (Block with the
    2x2 cross mask): PINNED -- ``tests/golden/*.npz`` were produced by importing
    the unmodified reference from ``/root/reference`` (script:
    ``tests/golden/make_golden.py``) and ``tests/test_oracle_golden.py`` checks
    every function here against them.
  * effective rank (``oracle/erank_oracle.py``): PARITY UNPINNED, the reference
    holds no effective-rank code at all.

All citations are relative to ``/root/reference``.
Shapes: rgb, depth are (B, T, C); the stacked output is (B, T, 2, C).
"""
from __future__ import annotations

import numpy as np

try:  # exact GELU needs erf; scipy is in the image, math.erf is the fallback
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    import math

    _erf = np.vectorize(math.erf, otypes=[np.float64])

BLEND_SWAP = 0    # model/futr_safuser_tokenfusion.py:59-60
BLEND_SCALE = 1   # model/futr_safuser_tokenfusion_vary.py:51-56
BLEND_CONVEX = 2  # model/futr_safuser_batchnormalization.py:65-74


# --------------------------------------------------------------------------
# a1: channel score
# --------------------------------------------------------------------------
def channel_score(x: np.ndarray) -> np.ndarray:
    """``x.abs().mean(dim=(0, 1))`` -- model/futr_safuser_tokenfusion.py:49-50
    (same expression in model/futr_safuser_tokenfusion_vary.py:41-42).

    Returns a (C,) float32 vector.  Accumulation is float64 then rounded once,
    so the oracle is at least as accurate as any fp32 summation order."""
    x = np.asarray(x)
    B, T, C = x.shape
    s = np.abs(x.astype(np.float64)).reshape(B * T, C).sum(axis=0)
    return (s / float(B * T)).astype(np.float32)


def train_branch_score(B: int, T: int, C: int) -> np.ndarray:
    """mode == 'train' branch -- model/futr_safuser_tokenfusion.py:40-45.

    ``autograd.grad(rgb.mean() + depth.mean(), [rgb, depth])`` is the constant
    1/(B*T*C) for every element, so ``abs().mean(dim=(0,1))`` is that same
    constant for every channel: the score is data independent and all C
    channels tie."""
    return np.full((C,), np.float32(1.0) / np.float32(B * T * C), dtype=np.float32)


# --------------------------------------------------------------------------
# a4: bottom-k
# --------------------------------------------------------------------------
def bottomk(score: np.ndarray, k: int) -> np.ndarray:
    """``torch.topk(score, k, dim=-1, largest=False)[1]`` --
    model/futr_safuser_tokenfusion.py:52-54.

    Ascending by score; ties go to the LOWER channel index (north_star rule).
    On tie-free scores this is exactly what ``torch.topk`` returns; on ties
    torch's order is an implementation artefact and this rule is
    the documented deviation."""
    score = np.asarray(score, dtype=np.float32).reshape(-1)
    order = np.argsort(score, kind="stable")
    return order[:k].astype(np.int64)


def k_for(variant: str, C: int) -> int:
    """k = C // 4 (tokenfusion.py:52, vary.py:44) or max(0, int(C * 0.1))
    (batchnormalization.py:58)."""
    if variant == "batchnorm":
        return max(0, int(C * 0.1))
    return C // 4


# --------------------------------------------------------------------------
# a3: BatchNorm1d front end of the BN variant
# --------------------------------------------------------------------------
def batchnorm_train(x, weight, bias, eps=1e-5):
    """``bn(x.permute(0,2,1)).permute(0,2,1)`` in training mode --
    model/futr_safuser_batchnormalization.py:45-46.

    Per channel over the B*T rows: biased variance for normalisation.  Returns
    (y, mean, biased_var, unbiased_var); running stats update with momentum 0.1
    uses the unbiased variance (torch.nn.BatchNorm1d semantics)."""
    x = np.asarray(x)
    B, T, C = x.shape
    xr = x.astype(np.float64).reshape(B * T, C)
    mean = xr.mean(axis=0)
    var_b = xr.var(axis=0)
    n = B * T
    var_u = var_b * n / max(n - 1, 1)
    y = (xr - mean) / np.sqrt(var_b + eps) * weight.astype(np.float64) + bias.astype(np.float64)
    return (y.reshape(B, T, C).astype(np.float32), mean.astype(np.float32),
            var_b.astype(np.float32), var_u.astype(np.float32))


def batchnorm_eval(x, weight, bias, running_mean, running_var, eps=1e-5):
    """Same call site with ``module.eval()``: running statistics."""
    x = np.asarray(x, dtype=np.float64)
    y = (x - running_mean.astype(np.float64)) / np.sqrt(running_var.astype(np.float64) + eps)
    y = y * weight.astype(np.float64) + bias.astype(np.float64)
    return y.astype(np.float32)


# --------------------------------------------------------------------------
# a5 / a6 / a7: exchange + stack
# --------------------------------------------------------------------------
def exchange_fwd(rgb, depth, idx_r, idx_d, alpha=None, blend=BLEND_SWAP):
    """clone + indexed assign + stack.

    swap   : model/futr_safuser_tokenfusion.py:56-62
    scale  : model/futr_safuser_tokenfusion_vary.py:48-57   (alpha * other)
    convex : model/futr_safuser_batchnormalization.py:62-75 (alpha*own + (1-alpha)*other)

    The right-hand sides read the ORIGINAL tensors, so a channel present in
    both index sets is exchanged in both directions."""
    rgb = np.asarray(rgb)
    depth = np.asarray(depth)
    ex_r = rgb.copy()
    ex_d = depth.copy()
    idx_r = np.asarray(idx_r, dtype=np.int64)
    idx_d = np.asarray(idx_d, dtype=np.int64)
    if blend == BLEND_SWAP:
        ex_r[:, :, idx_r] = depth[:, :, idx_r]
        ex_d[:, :, idx_d] = rgb[:, :, idx_d]
    elif blend == BLEND_SCALE:
        a = np.asarray(alpha, dtype=np.float32).reshape(-1)
        ex_r[:, :, idx_r] = (a[idx_r] * depth[:, :, idx_r].astype(np.float32)).astype(rgb.dtype)
        ex_d[:, :, idx_d] = (a[idx_d] * rgb[:, :, idx_d].astype(np.float32)).astype(rgb.dtype)
    elif blend == BLEND_CONVEX:
        a = np.asarray(alpha, dtype=np.float32).reshape(-1)
        r32 = rgb.astype(np.float32)
        d32 = depth.astype(np.float32)
        ex_r[:, :, idx_r] = (a[idx_r] * r32[:, :, idx_r] + (1 - a[idx_r]) * d32[:, :, idx_r]).astype(rgb.dtype)
        ex_d[:, :, idx_d] = (a[idx_d] * d32[:, :, idx_d] + (1 - a[idx_d]) * r32[:, :, idx_d]).astype(rgb.dtype)
    else:
        raise ValueError(blend)
    return np.stack([ex_r, ex_d], axis=2)


# --------------------------------------------------------------------------
# a8: backward of the exchange (verified there against
# the reference's autograd; re-verified here by tests/test_oracle_golden.py)
# --------------------------------------------------------------------------
def exchange_bwd(g, rgb, depth, idx_r, idx_d, alpha=None, blend=BLEND_SWAP):
    """Gradient of :func:`exchange_fwd` w.r.t. (rgb, depth, alpha).

    autograd of model/futr_safuser_tokenfusion.py:56-62 (and vary.py:48-57,
    batchnormalization.py:62-75).  It is a per-channel masked select, not a
    scatter-add: indices are unique within a modality."""
    g = np.asarray(g, dtype=np.float64)
    B, T, two, C = g.shape
    g_r, g_d = g[:, :, 0], g[:, :, 1]
    m_r = np.zeros(C); m_r[np.asarray(idx_r, dtype=np.int64)] = 1.0
    m_d = np.zeros(C); m_d[np.asarray(idx_d, dtype=np.int64)] = 1.0
    r = np.asarray(rgb, dtype=np.float64)
    d = np.asarray(depth, dtype=np.float64)
    if blend == BLEND_SWAP:
        d_rgb = g_r * (1 - m_r) + g_d * m_d
        d_dep = g_d * (1 - m_d) + g_r * m_r
        d_alpha = None
    elif blend == BLEND_SCALE:
        a = np.asarray(alpha, dtype=np.float64).reshape(-1)
        d_rgb = g_r * (1 - m_r) + a * g_d * m_d
        d_dep = g_d * (1 - m_d) + a * g_r * m_r
        d_alpha = m_r * (g_r * d).sum(axis=(0, 1)) + m_d * (g_d * r).sum(axis=(0, 1))
    elif blend == BLEND_CONVEX:
        a = np.asarray(alpha, dtype=np.float64).reshape(-1)
        d_rgb = g_r * ((1 - m_r) + a * m_r) + g_d * (1 - a) * m_d
        d_dep = g_d * ((1 - m_d) + a * m_d) + g_r * (1 - a) * m_r
        d_alpha = m_r * (g_r * (r - d)).sum(axis=(0, 1)) + m_d * (g_d * (d - r)).sum(axis=(0, 1))
    else:
        raise ValueError(blend)
    out = (d_rgb.astype(np.float32), d_dep.astype(np.float32),
           None if d_alpha is None else d_alpha.astype(np.float32))
    return out


# --------------------------------------------------------------------------
# token_fusion, per variant
# --------------------------------------------------------------------------
def token_fusion(variant, rgb, depth, mode, params=None, bn_training=True, return_indices=False):
    """``CMFuser.token_fusion(rgb, depth, mode)``.

    variant 'tokenfusion': model/futr_safuser_tokenfusion.py:33-66
    variant 'vary'       : model/futr_safuser_tokenfusion_vary.py:34-59
    variant 'batchnorm'  : model/futr_safuser_batchnormalization.py:38-77

    ``params`` holds numpy arrays under the reference state_dict names
    ('alpha', 'bn_rgb.weight', ...)."""
    rgb = np.asarray(rgb)
    depth = np.asarray(depth)
    B, T, C = rgb.shape
    k = k_for(variant, C)
    if variant == "tokenfusion":
        if mode == "train":
            s_r = train_branch_score(B, T, C)
            s_d = train_branch_score(B, T, C)
        else:
            s_r, s_d = channel_score(rgb), channel_score(depth)
        idx_r, idx_d = bottomk(s_r, k), bottomk(s_d, k)
        out = exchange_fwd(rgb, depth, idx_r, idx_d, None, BLEND_SWAP)
    elif variant == "vary":
        s_r, s_d = channel_score(rgb), channel_score(depth)
        idx_r, idx_d = bottomk(s_r, k), bottomk(s_d, k)
        out = exchange_fwd(rgb, depth, idx_r, idx_d, params["alpha"], BLEND_SCALE)
    elif variant == "batchnorm":
        if bn_training:
            rgb_n = batchnorm_train(rgb, params["bn_rgb.weight"], params["bn_rgb.bias"])[0]
            dep_n = batchnorm_train(depth, params["bn_depth.weight"], params["bn_depth.bias"])[0]
        else:
            rgb_n = batchnorm_eval(rgb, params["bn_rgb.weight"], params["bn_rgb.bias"],
                                   params["bn_rgb.running_mean"], params["bn_rgb.running_var"])
            dep_n = batchnorm_eval(depth, params["bn_depth.weight"], params["bn_depth.bias"],
                                   params["bn_depth.running_mean"], params["bn_depth.running_var"])
        s_r = np.abs(params["bn_rgb.weight"]).astype(np.float32)
        s_d = np.abs(params["bn_depth.weight"]).astype(np.float32)
        idx_r, idx_d = bottomk(s_r, k), bottomk(s_d, k)
        out = exchange_fwd(rgb_n, dep_n, idx_r, idx_d, params["alpha"], BLEND_CONVEX)
    else:
        raise ValueError(variant)
    if return_indices:
        return out, idx_r, idx_d
    return out


# --------------------------------------------------------------------------
# a9 / a10 / a11: the wrapper (Block with the 2x2 cross mask), eval semantics
# (Dropout is the identity; parity runs use .eval())
# --------------------------------------------------------------------------
def _layernorm(x, w, b, eps=1e-5):
    mu = x.mean(axis=-1, keepdims=True)
    var = x.var(axis=-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * w + b


def _gelu(x):
    return 0.5 * x * (1.0 + _erf(x / np.sqrt(2.0)))


def attention_2tok(x, qkv_w, proj_w, proj_b, num_heads, qkv_b=None):
    """``Attention.forward(x, attn_mask)`` with the mask of
    model/futr_safuser_tokenfusion.py:68-72 -- model/extras/transformerblock.py:19-36.

    Written out in full (q, k, softmax with -inf on the diagonal) so that the
    degenerate-attention shortcut the CUDA path takes is checked
    against the unsimplified arithmetic.  x: (R, 2, C) float64."""
    R, N, C = x.shape
    hd = C // num_heads
    qkv = x @ qkv_w.T
    if qkv_b is not None:
        qkv = qkv + qkv_b
    qkv = qkv.reshape(R, N, 3, num_heads, hd).transpose(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]                       # (R, H, N, hd)
    attn = (q @ k.transpose(0, 1, 3, 2)) * (hd ** -0.5)    # (R, H, N, N)
    mask = np.zeros((N, N)); mask[np.arange(N), np.arange(N)] = -np.inf
    attn = attn + mask
    attn = attn - attn.max(axis=-1, keepdims=True)
    e = np.exp(attn)
    attn = e / e.sum(axis=-1, keepdims=True)
    out = (attn @ v).transpose(0, 2, 1, 3).reshape(R, N, C)
    return out @ proj_w.T + proj_b, attn


def block_forward(x, p, prefix, num_heads):
    """``Block.forward`` -- model/extras/transformerblock.py:131-135."""
    g = lambda n: p[prefix + n].astype(np.float64)
    h = _layernorm(x, g("norm1.weight"), g("norm1.bias"))
    qkv_b = p.get(prefix + "attn.qkv.bias")
    a, attn = attention_2tok(h, g("attn.qkv.weight"), g("attn.proj.weight"), g("attn.proj.bias"),
                             num_heads, None if qkv_b is None else qkv_b.astype(np.float64))
    x = x + a
    h = _layernorm(x, g("norm2.weight"), g("norm2.bias"))
    h = _gelu(h @ g("mlp.mlp.0.weight").T + g("mlp.mlp.0.bias"))
    h = h @ g("mlp.mlp.2.weight").T + g("mlp.mlp.2.bias")
    return x + h, attn


def fuser_forward(variant, rgb, depth, mode, params, num_heads, depth_blocks=1, bn_training=False):
    """``CMFuser.forward(modal_feats, mode)`` in eval (Dropout = identity).

    tokenfusion: model/futr_safuser_tokenfusion.py:74-97 (outer residual :92)
    vary       : model/futr_safuser_tokenfusion_vary.py:67-87 (no outer residual)
    batchnorm  : model/futr_safuser_batchnormalization.py:85-107 (no outer residual)
    safuser    : model/futr_safuser_depth.py:37-64 (no exchange, + modality_token;
                 also returns the attention weights)"""
    rgb = np.asarray(rgb)
    B, T, C = rgb.shape
    if variant == "safuser":
        st = np.stack([rgb, depth], axis=2).astype(np.float64) + params["modality_token"].astype(np.float64)
    else:
        st = token_fusion(variant, rgb, depth, mode, params, bn_training=bn_training).astype(np.float64)
    x = st.reshape(B * T, 2, C)
    x_res = x
    attns = []
    for i in range(depth_blocks):
        x, attn = block_forward(x, params, f"blocks.{i}.", num_heads)
        attns.append(attn.reshape(B, T, *attn.shape[1:]))
    if variant == "tokenfusion":
        x = x + x_res
    x = _layernorm(x, params["norm.weight"].astype(np.float64), params["norm.bias"].astype(np.float64))
    y = x.mean(axis=1).reshape(B, T, C).astype(np.float32)
    if variant == "safuser":
        # torch.stack(attn_weights).transpose(0, 1): (B, depth, T, heads, 2, 2)
        return y, np.stack(attns).transpose(1, 0, 2, 3, 4, 5).astype(np.float32)
    return y


# ---------------------------------------------------------------------------------------------------------
# Token-axis selection (north_star kernels 3-6).  NO reference symbol: the reference ships the channel exchange only
#  and describes the token form in prose (README.md:13).  PARITY UNPINNED -- this restates the
# definition r3d_b200 fixes: per sample, the k = T // 4 tokens with the lowest spectral informativeness
# (oracle/erank_oracle.py:token_scores; ties -> lower token index) of a modality are replaced by the other modality's
# tokens at the same positions; output stacked (B, T, 2, C) like tokenfusion.py:62.
# ---------------------------------------------------------------------------------------------------------
def token_fusion_tokens(rgb, depth, k=None, rtol=1e-4, scores=None, return_indices=False):
    from . import erank_oracle as EO
    rgb, depth = np.asarray(rgb), np.asarray(depth)
    B, T, C = rgb.shape
    k = T // 4 if k is None else k
    s_r, s_d = scores if scores is not None else (EO.token_scores(rgb, rtol), EO.token_scores(depth, rtol))
    idx_r = np.stack([bottomk(np.asarray(s_r[b], np.float32), k) for b in range(B)]) if B else np.zeros((0, k), np.int64)
    idx_d = np.stack([bottomk(np.asarray(s_d[b], np.float32), k) for b in range(B)]) if B else np.zeros((0, k), np.int64)
    ex_r, ex_d = rgb.copy(), depth.copy()
    for b in range(B):
        ex_r[b, idx_r[b]] = depth[b, idx_r[b]]
        ex_d[b, idx_d[b]] = rgb[b, idx_d[b]]
    out = np.stack([ex_r, ex_d], axis=2)
    return (out, idx_r, idx_d) if return_indices else out


def token_exchange_bwd(g, idx_r, idx_d):
    """g (B, T, 2, C) -> (d_rgb, d_depth): masked select per sample (no duplicates inside an index set)."""
    g = np.asarray(g)
    B, T, _, C = g.shape
    g_r, g_d = g[:, :, 0], g[:, :, 1]
    m_r = np.zeros((B, T, 1), g.dtype); m_d = np.zeros((B, T, 1), g.dtype)
    for b in range(B):
        m_r[b, idx_r[b]] = 1
        m_d[b, idx_d[b]] = 1
    return g_r * (1 - m_r) + g_d * m_d, g_d * (1 - m_d) + g_r * m_r
