"""Host-side operators over the C ABI: thin, allocation + pointer plumbing only.

PyTorch supplies device memory and the current stream; every computation is a
hand-written sm_100a kernel in csrc/.  CPU tensors are rejected (no fallback).

Public surface (also registered as ``torch.ops.r3d.*``):
  channel_score, bottomk, exchange (autograd), token_fusion_bn (autograd),
  erank (autograd), gram, jacobi_eigh, token_informativeness.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import R3DError, check

BLEND_SWAP, BLEND_SCALE, BLEND_CONVEX = 0, 1, 2
DEFAULT_RTOL = 1e-4
_DT = {torch.float32: 0, torch.bfloat16: 1}


def _p(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_get_dev = getattr(torch._C, "_cuda_getDevice", None)
_xchg_dev = getattr(torch._C, "_cuda_exchangeDevice", None)
_maybe_xchg_dev = getattr(torch._C, "_cuda_maybeExchangeDevice", None)


def _stream():
    """Raw cudaStream_t of torch's current stream on the current device (the C accessors are ~20x cheaper than
    torch.cuda.current_stream(), which matters when a call is 8 launches of ~20 us kernels)."""
    if _raw_stream is not None and _get_dev is not None:
        return ctypes.c_void_p(_raw_stream(_get_dev()))
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class _on:
    """Device guard: make the tensor's device current for the launches inside (cheap when it already is)."""
    __slots__ = ("idx", "prev", "ctx")

    def __init__(self, device):
        self.idx = device.index if device.index is not None else torch.cuda.current_device()
        self.prev = -1
        self.ctx = None

    def __enter__(self):
        if _xchg_dev is not None and _maybe_xchg_dev is not None:
            self.prev = _xchg_dev(self.idx)
        else:
            self.ctx = torch.cuda.device(self.idx)
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            return self.ctx.__exit__(*exc)
        _maybe_xchg_dev(self.prev)
        return False


def _need_cuda(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise R3DError("r3d_b200 operators run on CUDA tensors only (no CPU fallback); got device "
                           f"{t.device}")


def _dt(t: torch.Tensor) -> int:
    if t.dtype not in _DT:
        raise R3DError(f"dtype must be float32 or bfloat16, got {t.dtype}")
    return _DT[t.dtype]


def _btc(rgb: torch.Tensor, depth: torch.Tensor):
    if rgb.dim() != 3 or depth.dim() != 3:
        raise R3DError(f"expected (B, T, C) tensors, got {tuple(rgb.shape)} and {tuple(depth.shape)}")
    if rgb.shape != depth.shape or rgb.dtype != depth.dtype or rgb.device != depth.device:
        raise R3DError("rgb and depth must agree in shape, dtype and device: "
                       f"{tuple(rgb.shape)} {rgb.dtype} {rgb.device} vs {tuple(depth.shape)} {depth.dtype} {depth.device}")
    return rgb.contiguous(), depth.contiguous()


# ---------------------------------------------------------------------------------
# a1: channel score            (reference: model/futr_safuser_tokenfusion.py:49-50)
# ---------------------------------------------------------------------------------
def channel_score_sums(rgb: torch.Tensor, depth: torch.Tensor) -> torch.Tensor:
    """Column sums of |x| for both modalities, (2, C) float32 (not divided by rows)."""
    _need_cuda(rgb, depth)
    rgb, depth = _btc(rgb, depth)
    B, T, C = rgb.shape
    rows = B * T
    L = _lib.lib()
    with _on(rgb.device):
        ws = torch.empty(L.r3d_score_workspace_floats(rows, C), dtype=torch.float32, device=rgb.device)
        sums = torch.empty(2, C, dtype=torch.float32, device=rgb.device)
        check(L.r3d_channel_score_partial(_p(rgb), _p(depth), rows, C, _dt(rgb), _p(ws), _stream()))
        check(L.r3d_score_finalize(_p(ws), max(rows, 1), C, _p(sums), None, _stream()))
    return sums


def channel_score(rgb: torch.Tensor, depth: torch.Tensor) -> torch.Tensor:
    """(2, C) float32: row 0 = rgb.abs().mean((0,1)), row 1 = depth.abs().mean((0,1))."""
    _need_cuda(rgb, depth)
    rgb, depth = _btc(rgb, depth)
    B, T, C = rgb.shape
    rows = B * T
    L = _lib.lib()
    with _on(rgb.device):
        ws = torch.empty(L.r3d_score_workspace_floats(rows, C), dtype=torch.float32, device=rgb.device)
        score = torch.empty(2, C, dtype=torch.float32, device=rgb.device)
        check(L.r3d_channel_score_partial(_p(rgb), _p(depth), rows, C, _dt(rgb), _p(ws), _stream()))
        check(L.r3d_score_finalize(_p(ws), max(rows, 1), C, None, _p(score), _stream()))
    return score


def channel_score_packed(rgb: torch.Tensor, depth: torch.Tensor, er: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(2C + 2,) float32 packed statistic [sum|rgb| (C) | sum|depth| (C) | sum er | rows], written by one finalize
    launch -- the buffer global-score mode all-reduces; `bottomk_packed` consumes it."""
    _need_cuda(rgb, depth, er)
    rgb, depth = _btc(rgb, depth)
    B, T, C = rgb.shape
    rows = B * T
    L = _lib.lib()
    with _on(rgb.device):
        ws = torch.empty(L.r3d_score_workspace_floats(rows, C), dtype=torch.float32, device=rgb.device)
        packed = torch.empty(2 * C + 2, dtype=torch.float32, device=rgb.device)
        check(L.r3d_channel_score_partial(_p(rgb), _p(depth), rows, C, _dt(rgb), _p(ws), _stream()))
        e = None if er is None else er.reshape(-1).float().contiguous()
        check(L.r3d_score_finalize_packed(_p(ws), max(rows, 1), C, _p(e), 0 if e is None else e.numel(), _p(packed),
                                          _stream()))
    return packed


def bottomk_packed(packed: torch.Tensor, k: int, return_score: bool = False):
    """Bottom-k of  sums / rows  straight from the packed statistic: (2, k) int64 (+ the (2, C) score)."""
    _need_cuda(packed)
    C = (packed.numel() - 2) // 2
    idx = torch.empty(2, k, dtype=torch.int64, device=packed.device)
    score = torch.empty(2, C, dtype=torch.float32, device=packed.device) if return_score else None
    with _on(packed.device):
        check(_lib.lib().r3d_bottomk_scaled(_p(packed), 2, C, k, packed.data_ptr() + (2 * C + 1) * 4, _p(idx),
                                            _p(score), _stream()))
    return (idx, score) if return_score else idx


# ---------------------------------------------------------------------------------
# a4: bottom-k                 (reference: model/futr_safuser_tokenfusion.py:52-54)
# ---------------------------------------------------------------------------------
def bottomk(score: torch.Tensor, k: int) -> torch.Tensor:
    """score (..., C) float32 -> (..., k) int64: ascending score, ties -> lower index."""
    _need_cuda(score)
    if score.dtype != torch.float32:
        raise R3DError("bottomk scores must be float32")
    C = score.shape[-1]
    s2 = score.reshape(-1, C).contiguous()
    out = torch.empty(s2.shape[0], k, dtype=torch.int64, device=score.device)
    with _on(score.device):
        check(_lib.lib().r3d_bottomk(_p(s2), s2.shape[0], C, k, _p(out), _stream()))
    return out.reshape(*score.shape[:-1], k)


# ---------------------------------------------------------------------------------
# a5-a8: exchange / blend with autograd
# ---------------------------------------------------------------------------------
def _exchange_fwd_raw(rgb, depth, idx_r, idx_d, alpha, affine, blend):
    B, T, C = rgb.shape
    out = torch.empty(B, T, 2, C, dtype=rgb.dtype, device=rgb.device)
    k = idx_r.numel()
    if B * T == 0:
        return out                                   # empty batch: nothing to launch (zero-size tensors have no pointer)
    with _on(rgb.device):
        check(_lib.lib().r3d_exchange_fwd(_p(rgb), _p(depth), _p(idx_r), _p(idx_d), k, _p(alpha), _p(affine), blend,
                                          _p(out), B * T, C, _dt(rgb), _stream()))
    return out


def _exchange_bwd_raw(g, rgb, depth, idx_r, idx_d, alpha, affine, bn_norm, blend):
    B, T, two, C = g.shape
    rows = B * T
    L = _lib.lib()
    d_rgb = torch.empty(B, T, C, dtype=g.dtype, device=g.device)
    d_dep = torch.empty(B, T, C, dtype=g.dtype, device=g.device)
    colsums = None
    if rows == 0:                                    # empty batch: zero parameter gradients, nothing to launch
        return d_rgb, d_dep, (None if blend == BLEND_SWAP else torch.zeros(5, C, dtype=torch.float32, device=g.device))
    with _on(g.device):
        ws = None
        if blend != BLEND_SWAP:
            ws = torch.empty(L.r3d_exchange_bwd_workspace_floats(rows, C), dtype=torch.float32, device=g.device)
        check(L.r3d_exchange_bwd(_p(g), _p(rgb), _p(depth), _p(idx_r), _p(idx_d), idx_r.numel(), _p(alpha),
                                 _p(affine), _p(bn_norm), blend, _p(d_rgb), _p(d_dep), _p(ws), rows, C, _dt(g),
                                 _stream()))
        if blend != BLEND_SWAP:
            colsums = torch.zeros(5, C, dtype=torch.float32, device=g.device)
            if bn_norm is None:
                # only the d_alpha row was written by the kernel
                ws[L.r3d_exchange_bwd_workspace_floats(rows, C) // 5:].zero_()
            check(L.r3d_exchange_bwd_finalize(_p(ws), max(rows, 1), C, _p(colsums), _stream()))
    return d_rgb, d_dep, colsums


class _Exchange(torch.autograd.Function):
    """exchange + stack (swap / alpha-scaled / convex), no BatchNorm front end."""

    @staticmethod
    def forward(ctx, rgb, depth, idx_r, idx_d, alpha, blend):
        ctx.blend = blend
        a = None if alpha is None else alpha.detach().reshape(-1).float().contiguous()
        out = _exchange_fwd_raw(rgb, depth, idx_r, idx_d, a, None, blend)
        if blend == BLEND_SWAP:
            ctx.save_for_backward(idx_r, idx_d)
        else:
            ctx.save_for_backward(idx_r, idx_d, rgb, depth, a)
        ctx.alpha_shape = None if alpha is None else alpha.shape
        ctx.alpha_dtype = None if alpha is None else alpha.dtype
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        if ctx.blend == BLEND_SWAP:
            idx_r, idx_d = ctx.saved_tensors
            d_rgb, d_dep, _ = _exchange_bwd_raw(g, None, None, idx_r, idx_d, None, None, None, BLEND_SWAP)
            return d_rgb, d_dep, None, None, None, None
        idx_r, idx_d, rgb, depth, a = ctx.saved_tensors
        d_rgb, d_dep, colsums = _exchange_bwd_raw(g, rgb, depth, idx_r, idx_d, a, None, None, ctx.blend)
        d_alpha = colsums[0].reshape(ctx.alpha_shape).to(ctx.alpha_dtype)
        return d_rgb, d_dep, None, None, d_alpha, None


def exchange(rgb, depth, idx_r, idx_d, alpha=None, blend=BLEND_SWAP) -> torch.Tensor:
    """(B,T,C) x2 + index sets -> stacked (B,T,2,C); differentiable in rgb, depth, alpha."""
    _need_cuda(rgb, depth, idx_r, idx_d, alpha)
    rgb, depth = _btc(rgb, depth)
    _dt(rgb)
    if blend != BLEND_SWAP and alpha is None:
        raise R3DError("alpha is required for the scale / convex blends")
    idx_r = idx_r.reshape(-1).to(torch.int64).contiguous()
    idx_d = idx_d.reshape(-1).to(torch.int64).contiguous()
    if idx_r.numel() != idx_d.numel():
        raise R3DError("idx_r and idx_d must have the same length")
    return _Exchange.apply(rgb, depth, idx_r, idx_d, alpha, blend)


# ---------------------------------------------------------------------------------
# a3 + a7: BatchNorm front end fused with the convex blend
# (reference: model/futr_safuser_batchnormalization.py:45-46, 62-75)
# ---------------------------------------------------------------------------------
def bn_batch_stats(rgb, depth) -> torch.Tensor:
    """(2 modalities, 3, C) float32: [mean, biased var, unbiased var] over the B*T rows."""
    _need_cuda(rgb, depth)
    rgb, depth = _btc(rgb, depth)
    B, T, C = rgb.shape
    L = _lib.lib()
    with _on(rgb.device):
        ws = torch.empty(L.r3d_bn_workspace_floats(B * T, C), dtype=torch.float32, device=rgb.device)
        stats = torch.empty(2, 3, C, dtype=torch.float32, device=rgb.device)
        check(L.r3d_bn_stats(_p(rgb), _p(depth), B * T, C, _dt(rgb), _p(ws), _p(stats), _stream()))
    return stats


class _TokenFusionBN(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rgb, depth, alpha, w_r, b_r, w_d, b_d, mean, var, idx_r, idx_d, eps, batch_stats):
        # mean/var: (2, C) statistics actually used for normalisation
        rstd = torch.rsqrt(var + eps)
        w = torch.stack([w_r.detach().float(), w_d.detach().float()])
        b = torch.stack([b_r.detach().float(), b_d.detach().float()])
        scale = w * rstd
        affine = torch.stack([scale, b - mean * scale], dim=1).contiguous()        # (2, 2, C)
        bn_norm = torch.stack([rstd, -mean * rstd], dim=1).contiguous()            # (2, 2, C)
        a = alpha.detach().reshape(-1).float().contiguous()
        out = _exchange_fwd_raw(rgb, depth, idx_r, idx_d, a, affine, BLEND_CONVEX)
        ctx.save_for_backward(rgb, depth, a, affine, bn_norm, idx_r, idx_d, w)
        ctx.batch_stats = batch_stats
        ctx.alpha_shape = alpha.shape
        return out

    @staticmethod
    def backward(ctx, g):
        rgb, depth, a, affine, bn_norm, idx_r, idx_d, w = ctx.saved_tensors
        g = g.contiguous()
        d_rgb, d_dep, cs = _exchange_bwd_raw(g, rgb, depth, idx_r, idx_d, a, affine, bn_norm, BLEND_CONVEX)
        d_alpha = cs[0].reshape(ctx.alpha_shape)
        d_w_r, d_b_r, d_w_d, d_b_d = cs[2], cs[1], cs[4], cs[3]
        B, T, C = rgb.shape
        sums = cs if ctx.batch_stats else torch.zeros_like(cs)   # running stats: no dependence on the batch
        with _on(g.device):
            check(_lib.lib().r3d_bn_bwd_apply(_p(rgb), _p(depth), _p(bn_norm), _p(w[0].contiguous()),
                                              _p(w[1].contiguous()), _p(sums), _p(d_rgb), _p(d_dep), B * T, C,
                                              _dt(rgb), _stream()))
        return d_rgb, d_dep, d_alpha, d_w_r, d_b_r, d_w_d, d_b_d, None, None, None, None, None, None


def token_fusion_bn(rgb, depth, alpha, w_r, b_r, w_d, b_d, mean, var, idx_r, idx_d, eps=1e-5, batch_stats=True):
    _need_cuda(rgb, depth, alpha, w_r, w_d, mean, var)
    rgb, depth = _btc(rgb, depth)
    return _TokenFusionBN.apply(rgb, depth, alpha, w_r, b_r, w_d, b_d, mean.float(), var.float(),
                                idx_r.reshape(-1).contiguous(), idx_d.reshape(-1).contiguous(), eps, batch_stats)


# ---------------------------------------------------------------------------------
# a12: effective rank (no reference symbol)
# ---------------------------------------------------------------------------------
GRAM_TCGEN05, GRAM_SIMT = 0, 1


def _erank_fwd_raw(x, rtol, gram_impl):
    B, T, C = x.shape
    n, m = min(T, C), max(T, C)
    L = _lib.lib()
    dev = x.device
    with _on(dev):
        ws = torch.empty(L.r3d_erank_workspace_bytes(B, T, C, _dt(x)), dtype=torch.uint8, device=dev)
        er = torch.empty(B, dtype=torch.float32, device=dev)
        sigma = torch.empty(B, n, dtype=torch.float32, device=dev)
        U = torch.empty(B, n, n, dtype=torch.float32, device=dev)
        Y = torch.empty(B, n, m, dtype=torch.float32, device=dev)
        sweeps = torch.empty(B, dtype=torch.int32, device=dev)
        check(L.r3d_erank_fwd(_p(x), B, T, C, _dt(x), rtol, gram_impl, _p(ws), _p(er), _p(sigma), _p(U), _p(Y),
                              _p(sweeps), _stream()))
    return er, sigma, U, Y, sweeps


class _ERank(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, rtol, gram_impl):
        er, sigma, U, Y, sweeps = _erank_fwd_raw(x, rtol, gram_impl)
        ctx.save_for_backward(er, sigma, U, Y)
        ctx.shape = x.shape
        ctx.dtype = x.dtype
        ctx.rtol = rtol
        ctx.mark_non_differentiable(sigma, sweeps)
        return er, sigma, sweeps

    @staticmethod
    def backward(ctx, g, _gs, _gw):
        er, sigma, U, Y = ctx.saved_tensors
        B, T, C = ctx.shape
        L = _lib.lib()
        dev = er.device
        dt = _DT[ctx.dtype]
        with _on(dev):
            ws = torch.empty(L.r3d_erank_workspace_bytes(B, T, C, dt), dtype=torch.uint8, device=dev)
            dx = torch.empty(B, T, C, dtype=ctx.dtype, device=dev)
            gg = g.contiguous().float()
            check(L.r3d_erank_bwd(_p(gg), _p(er), _p(sigma), _p(U), _p(Y), B, T, C, dt, ctx.rtol, _p(ws), _p(dx), 0,
                                  _stream()))
        return dx, None, None


def erank(x: torch.Tensor, rtol: float = DEFAULT_RTOL, gram_impl: int = GRAM_TCGEN05, return_aux: bool = False,
          strict: Optional[bool] = None):
    """Per-sample effective rank of x (B, T, C) -> (B,) float32, differentiable in x.

    exp(-sum p ln p), p = sigma / sum sigma over the singular values of each (T, C)
    sample; sigma <= rtol * sigma_max are treated as zero.

    ``return_aux`` adds (sigma, sweeps); sweeps < 0 marks a sample whose eigensolver ran into its sweep cap.
    ``strict=True`` (or environment R3D_STRICT=1) synchronises and raises R3DError in that case instead of returning a
    less accurate value silently; the default stays asynchronous."""
    _need_cuda(x)
    if x.dim() != 3:
        raise R3DError(f"expected (B, T, C), got {tuple(x.shape)}")
    _dt(x)
    er, sigma, sweeps = _ERank.apply(x.contiguous(), float(rtol), int(gram_impl))
    if strict is None:
        strict = os.environ.get("R3D_STRICT", "0") not in ("", "0")
    if strict and bool((sweeps < 0).any()):
        bad = torch.nonzero(sweeps < 0).flatten().tolist()
        raise R3DError(f"erank: the Jacobi eigensolver hit its sweep cap without converging for samples {bad[:8]} "
                       "(raise erank_pass2_sweeps / jacobi_max_sweeps via r3d_b200._lib.set_option)")
    if return_aux:
        return er, sigma, sweeps
    return er


def gram(x: torch.Tensor, gram_impl: int = GRAM_TCGEN05) -> torch.Tensor:
    """(B, T, C) -> (B, n, n) float32 Gram on the smaller side (X X^T if T <= C else X^T X)."""
    _need_cuda(x)
    x = x.contiguous()
    B, T, C = x.shape
    n = min(T, C)
    L = _lib.lib()
    with _on(x.device):
        ws = torch.empty(L.r3d_erank_workspace_bytes(B, T, C, _dt(x)), dtype=torch.uint8, device=x.device)
        G = torch.empty(B, n, n, dtype=torch.float32, device=x.device)
        check(L.r3d_gram(_p(x), B, T, C, _dt(x), gram_impl, _p(ws), _p(G), _stream()))
    return G


def jacobi_eigh(G: torch.Tensor, max_sweeps: int = 30) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Symmetric PSD (B, n, n) float32 -> (lambda (B, n), U (B, n, n), sweeps (B,)).  Solver order."""
    _need_cuda(G)
    if G.dtype != torch.float32 or G.dim() != 3 or G.shape[1] != G.shape[2]:
        raise R3DError("jacobi_eigh expects (B, n, n) float32")
    G = G.contiguous()
    B, n, _ = G.shape
    L = _lib.lib()
    with _on(G.device):
        ws = torch.empty(L.r3d_jacobi_workspace_bytes(B, n), dtype=torch.uint8, device=G.device)
        lam = torch.empty(B, n, dtype=torch.float32, device=G.device)
        U = torch.empty(B, n, n, dtype=torch.float32, device=G.device)
        sw = torch.empty(B, dtype=torch.int32, device=G.device)
        check(L.r3d_jacobi_eigh(_p(G), B, n, _p(ws), _p(lam), _p(U), _p(sw), max_sweeps, _stream()))
    return lam, U, sw


# ---------------------------------------------------------------------------------
# f1 (first step): row LayerNorm of the fuser Block
# (reference: model/extras/transformerblock.py:122,127,132,134; model/futr_safuser_tokenfusion.py:25,93)
# ---------------------------------------------------------------------------------
def layer_norm_supported(x: torch.Tensor, C: int) -> bool:
    """Shapes the hand-written kernel takes; anything else is the caller's business (CMFuser falls back to
    torch's LayerNorm -- library code on the same device, not a CPU path -- for widths it does not cover)."""
    if not x.is_cuda or x.dtype not in _DT:
        return False
    v = 4 if x.dtype == torch.float32 else 8
    # 128-bit row accesses: an offset (unaligned) view takes the caller's torch path instead of raising
    return C % v == 0 and C <= 32 * v * 8 and x.data_ptr() % 16 == 0 and x.is_contiguous()


class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps, pair_mean):
        C = x.shape[-1]
        xc = x.contiguous()
        rows = xc.numel() // C
        w = weight.to(xc.dtype).contiguous()
        b = bias.to(xc.dtype).contiguous()
        L = _lib.lib()
        with _on(xc.device):
            if pair_mean:
                y = torch.empty(xc.shape[:-2] + (C,), dtype=xc.dtype, device=xc.device)
            else:
                y = torch.empty_like(xc)
            mean = torch.empty(rows, dtype=torch.float32, device=xc.device)
            rstd = torch.empty(rows, dtype=torch.float32, device=xc.device)
            check(L.r3d_ln_fwd(_p(xc), _p(w), _p(b), rows, C, _dt(xc), float(eps), int(pair_mean), _p(y), _p(mean),
                               _p(rstd), _stream()))
        ctx.save_for_backward(xc, w, mean, rstd)
        ctx.wdtype, ctx.bdtype, ctx.pair = weight.dtype, bias.dtype, bool(pair_mean)
        return y

    @staticmethod
    def backward(ctx, gy):
        xc, w, mean, rstd = ctx.saved_tensors
        C = xc.shape[-1]
        rows = xc.numel() // C
        g = gy.contiguous()
        L = _lib.lib()
        with _on(xc.device):
            dx = torch.empty_like(xc)
            ws = torch.empty(L.r3d_ln_bwd_workspace_floats(rows, C), dtype=torch.float32, device=xc.device)
            dgb = torch.empty(2, C, dtype=torch.float32, device=xc.device)
            check(L.r3d_ln_bwd(_p(g), _p(xc), _p(mean), _p(rstd), _p(w), rows, C, _dt(xc), int(ctx.pair), _p(dx), _p(ws),
                               _p(dgb), _stream()))
        return dx, dgb[0].to(ctx.wdtype), dgb[1].to(ctx.bdtype), None, None


def _ln_args(x, weight, bias):
    _need_cuda(x, weight, bias)
    C = x.shape[-1]
    if weight.shape != (C,) or bias.shape != (C,):
        raise R3DError(f"layer_norm: weight/bias must have shape ({C},)")
    if not layer_norm_supported(x, C):
        raise R3DError(f"layer_norm: unsupported dtype/width {x.dtype}, C={C} (see include/r3d_b200.h)")


def layer_norm(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """F.layer_norm(x, (C,), weight, bias, eps) over the last dimension, differentiable in x, weight and bias."""
    _ln_args(x, weight, bias)
    return _LayerNorm.apply(x, weight, bias, eps, False)


def layer_norm_mean2(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """x (..., 2, C) -> (..., C): F.layer_norm over C followed by the mean over the two tokens, in one kernel
    (reference: model/futr_safuser_tokenfusion.py:93-95)."""
    _ln_args(x, weight, bias)
    if x.dim() < 2 or x.shape[-2] != 2:
        raise R3DError(f"layer_norm_mean2 expects (..., 2, C), got {tuple(x.shape)}")
    return _LayerNorm.apply(x, weight, bias, eps, True)


class _SwapAdd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p):
        xc, pc = x.contiguous(), p.contiguous()
        C = xc.shape[-1]
        rows = xc.numel() // C
        out = torch.empty_like(xc)
        with _on(xc.device):
            check(_lib.lib().r3d_swap_add(_p(xc), _p(pc), rows, C, _dt(xc), _p(out), _stream()))
        return out

    @staticmethod
    def backward(ctx, g):
        gc = g.contiguous()
        C = gc.shape[-1]
        rows = gc.numel() // C
        dp = torch.empty_like(gc)
        with _on(gc.device):
            check(_lib.lib().r3d_swap_add(None, _p(gc), rows, C, _dt(gc), _p(dp), _stream()))
        return g, dp


def swap_add_supported(x: torch.Tensor) -> bool:
    return x.is_cuda and x.dtype in _DT and x.dim() >= 2 and x.shape[-2] == 2 and \
        x.shape[-1] % (4 if x.dtype == torch.float32 else 8) == 0 and x.data_ptr() % 16 == 0


def swap_add(x: torch.Tensor, p: torch.Tensor) -> torch.Tensor:
    """x + p.flip(-2) for (..., 2, C) tensors: the residual add of the closed-form 2-token attention, where token m
    receives the projected V of token 1-m."""
    _need_cuda(x, p)
    if x.shape != p.shape or x.dtype != p.dtype or not swap_add_supported(x):
        raise R3DError(f"swap_add: expected two (..., 2, C) CUDA tensors of equal shape/dtype, got {tuple(x.shape)} {x.dtype}"
                       f" and {tuple(p.shape)} {p.dtype}")
    return _SwapAdd.apply(x, p)


def token_informativeness(sigma: torch.Tensor, U: torch.Tensor, rtol: float = DEFAULT_RTOL) -> torch.Tensor:
    """s_t = sum_j p_j U[t, j]^2 over the short side (B, n)."""
    _need_cuda(sigma, U)
    B, n = sigma.shape
    out = torch.empty(B, n, dtype=torch.float32, device=sigma.device)
    with _on(sigma.device):
        check(_lib.lib().r3d_token_informativeness(_p(sigma.contiguous()), _p(U.contiguous()), B, n, rtol, _p(out),
                                                   _stream()))
    return out


# ---------------------------------------------------------------------------------
# f1 / f2 / f3: tcgen05 linear GEMM with fused epilogues (csrc/linear_tcgen05.cu)
# ---------------------------------------------------------------------------------
ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2


def gemm(A: torch.Tensor, B: torch.Tensor, a_kmajor: bool = True, b_kmajor: bool = True, *, bias=None, act: int = 0,
         residual=None, want_aux: bool = False, aux_in=None, colsum: bool = False, colsum_abs: bool = False, out=None):
    """D (M, N) = epilogue(sum_k A[m, k] B[n, k]) on the hand-written tcgen05 GEMM (no library call).

    a_kmajor: A is (M, K), else (K, M);  b_kmajor: B is (N, K) [an nn.Linear weight], else (K, N).
    Epilogue: +bias -> aux (pre-activation copy) -> act -> * gelu'(aux_in) -> +residual -> D; optional column sums of D
    (returned as (N,) float32 after the fixed-order finalize).  Returns D, or (D, aux, colsum) items as requested."""
    _need_cuda(A, B, bias, residual, aux_in)
    dt = _dt(A)
    if B.dtype != A.dtype:
        raise R3DError("gemm operands must have the same dtype")
    A, B = A.contiguous(), B.contiguous()
    M, K = (A.shape if a_kmajor else A.shape[::-1])
    N, K2 = (B.shape if b_kmajor else B.shape[::-1])
    if K != K2:
        raise R3DError(f"gemm: contraction sizes differ ({K} vs {K2})")
    L = _lib.lib()
    dev = A.device
    with _on(dev):
        D = torch.empty(M, N, dtype=A.dtype, device=dev) if out is None else out
        ws = torch.empty(L.r3d_gemm_workspace_bytes(M, N, K, dt), dtype=torch.uint8, device=dev)
        aux = torch.empty(M, N, dtype=A.dtype, device=dev) if want_aux else None
        parts = (M + 127) // 128
        cs = torch.empty(parts, N, dtype=torch.float32, device=dev) if colsum else None
        ep = _lib.Epilogue()
        ep.bias = None if bias is None else bias.to(A.dtype).contiguous().data_ptr()
        ep.residual = None if residual is None else residual.contiguous().data_ptr()
        ep.aux_out = None if aux is None else aux.data_ptr()
        ep.aux_in = None if aux_in is None else aux_in.contiguous().data_ptr()
        ep.colsum_partial = None if cs is None else cs.data_ptr()
        ep.act, ep.colsum_abs = int(act), int(bool(colsum_abs))
        check(L.r3d_gemm(_p(A), _p(B), _p(D), M, N, K, int(a_kmajor), int(b_kmajor), dt, ctypes.byref(ep), _p(ws),
                         _stream()))
        csum = None
        if colsum:
            csum = torch.empty(N, dtype=torch.float32, device=dev)
            check(L.r3d_colsum_finalize(_p(cs), parts, N, 0, _p(csum), _stream()))
    outs = [D]
    if want_aux:
        outs.append(aux)
    if colsum:
        outs.append(csum)
    return outs[0] if len(outs) == 1 else tuple(outs)


def colsum(x: torch.Tensor) -> torch.Tensor:
    """Column sums of a (rows, C) tensor in its own dtype (deterministic two-stage tree): a bias gradient."""
    _need_cuda(x)
    x = x.contiguous()
    rows, C = x.shape
    L = _lib.lib()
    with _on(x.device):
        ws = torch.empty(L.r3d_colsum_workspace_floats(rows, C), dtype=torch.float32, device=x.device)
        out = torch.empty(C, dtype=x.dtype, device=x.device)
        check(L.r3d_colsum(_p(x), rows, C, _dt(x), _p(ws), _p(out), _stream()))
    return out


# ---------------------------------------------------------------------------------
# f1: the whole fuser Block (transformerblock.py:118-135 with the 2-token mask of tokenfusion.py:68-72, closed form of
# SURVEY.md F4) as ONE autograd node over hand-written kernels only: 2 LayerNorm + 4 GEMM launches forward,
# 2 LayerNorm + 8 GEMM + 2 column-sum launches backward; bias / GELU / residual / gelu' / bias-gradient sums live in the
# GEMM epilogues, the token swap and the residual-gradient add in the LayerNorm kernels.
# ---------------------------------------------------------------------------------
def _ln_fwd_raw(x2d, w, b, eps, flags):
    rows, C = x2d.shape
    L = _lib.lib()
    y = torch.empty_like(x2d)
    mean = torch.empty(rows, dtype=torch.float32, device=x2d.device)
    rstd = torch.empty(rows, dtype=torch.float32, device=x2d.device)
    check(L.r3d_ln_fwd2(_p(x2d), _p(w), _p(b), rows, C, _dt(x2d), float(eps), int(flags), _p(y), _p(mean), _p(rstd),
                        _stream()))
    return y, mean, rstd


def _ln_bwd_raw(dy, x2d, mean, rstd, w, flags, addend):
    rows, C = x2d.shape
    L = _lib.lib()
    dx = torch.empty_like(x2d)
    ws = torch.empty(L.r3d_ln_bwd_workspace_floats(rows, C), dtype=torch.float32, device=x2d.device)
    dgb = torch.empty(2, C, dtype=torch.float32, device=x2d.device)
    check(L.r3d_ln_bwd2(_p(dy), _p(x2d), _p(mean), _p(rstd), _p(w), rows, C, _dt(x2d), int(flags), _p(addend), _p(dx),
                        _p(ws), _p(dgb), _stream()))
    return dx, dgb[0], dgb[1]


def fused_block_supported(x: torch.Tensor, C: int, hidden: int) -> bool:
    return (x.is_cuda and x.dtype in _DT and x.dim() == 3 and x.shape[1] == 2 and C % 8 == 0 and hidden % 8 == 0
            and layer_norm_supported(x, C))


class _FusedBlock(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, n1w, n1b, qkv_w, qkv_b, wp, bp, n2w, n2b, w1, b1, w2, b2, eps1, eps2):
        dt = x.dtype
        C = x.shape[-1]
        xc = x.contiguous().view(-1, C)
        cast = lambda t: None if t is None else t.detach().to(dt).contiguous()
        n1w_, n1b_, n2w_, n2b_ = cast(n1w), cast(n1b), cast(n2w), cast(n2b)
        wv = cast(qkv_w[2 * C:])
        bv = None if qkv_b is None else cast(qkv_b[2 * C:])
        wp_, bp_, w1_, b1_, w2_, b2_ = cast(wp), cast(bp), cast(w1), cast(b1), cast(w2), cast(b2)
        with _on(x.device):
            h1sw, m1, r1 = _ln_fwd_raw(xc, n1w_, n1b_, eps1, 2)          # norm1, rows of each pair swapped on the write
            vsw = gemm(h1sw, wv, bias=bv)                                 # V of the OTHER token (qkv's V third only)
            x1 = gemm(vsw, wp_, bias=bp_, residual=xc)                    # x + proj(V[other])
            h2, m2, r2 = _ln_fwd_raw(x1, n2w_, n2b_, eps2, 0)
            g1, H = gemm(h2, w1_, bias=b1_, act=ACT_GELU, want_aux=True)  # GELU(fc1) with the pre-activation saved
            x2 = gemm(g1, w2_, bias=b2_, residual=x1)                     # x1 + fc2
        ctx.save_for_backward(xc, m1, r1, h1sw, vsw, x1, m2, r2, h2, H, g1, n1w_, wv, wp_, n2w_, w1_, w2_)
        ctx.meta = (x.shape, qkv_w.shape, qkv_b is not None,
                    [None if t is None else t.dtype for t in (n1w, n1b, qkv_w, qkv_b, wp, bp, n2w, n2b, w1, b1, w2, b2)])
        return x2.view(x.shape)

    @staticmethod
    def backward(ctx, dx2):
        xc, m1, r1, h1sw, vsw, x1, m2, r2, h2, H, g1, n1w, wv, wp, n2w, w1, w2 = ctx.saved_tensors
        xshape, qkv_shape, has_qkv_b, pdt = ctx.meta
        C = xc.shape[1]
        dx2 = dx2.contiguous().view(-1, C)
        with _on(dx2.device):
            db2 = colsum(dx2)
            dH, db1 = gemm(dx2, w2, True, False, aux_in=H, colsum=True)      # (dx2 W2) * gelu'(H), + column sums = d b1
            dW2 = gemm(dx2, g1, False, False)
            dh2 = gemm(dH, w1, True, False)
            dW1 = gemm(dH, h2, False, False)
            dx1, dg2, dbt2 = _ln_bwd_raw(dh2, x1, m2, r2, n2w, 0, dx2)       # norm2 backward + residual gradient
            dbp = colsum(dx1)
            dvsw = gemm(dx1, wp, True, False)
            dWp = gemm(dx1, vsw, False, False)
            dh1sw = gemm(dvsw, wv, True, False)
            dWv = gemm(dvsw, h1sw, False, False)
            dx, dg1, dbt1 = _ln_bwd_raw(dh1sw, xc, m1, r1, n1w, 2, dx1)      # norm1 backward (dy read swapped) + residual
            dqkv = torch.zeros(qkv_shape, dtype=dWv.dtype, device=dWv.device)  # W_q / W_k: exactly zero
            dqkv[2 * C:] = dWv
            dqkv_b = None
            if has_qkv_b:
                dqkv_b = torch.zeros(qkv_shape[0], dtype=dWv.dtype, device=dWv.device)
                dqkv_b[2 * C:] = colsum(dvsw)
        g = [dg1, dbt1, dqkv, dqkv_b, dWp, dbp, dg2, dbt2, dW1, db1, dW2, db2]
        g = [None if (t is None or d is None) else t.to(d) for t, d in zip(g, pdt)]
        return (dx.view(xshape), *g, None, None)


def fused_block(x, norm1, qkv, proj, norm2, fc1, fc2):
    """x (R, 2, C) -> (R, 2, C): LN -> V -> proj(+swapped residual) -> LN -> MLP(+residual), see _FusedBlock."""
    return _FusedBlock.apply(x, norm1.weight, norm1.bias, qkv.weight, qkv.bias, proj.weight, proj.bias, norm2.weight,
                             norm2.bias, fc1.weight, fc1.bias, fc2.weight, fc2.bias, norm1.eps, norm2.eps)


# ---------------------------------------------------------------------------------
# N2: M-modality fuser pieces (csrc/multi_kernels.cu): cyclic channel exchange, M-token masked attention, token mean
# ---------------------------------------------------------------------------------
class _ExchangeMulti(torch.autograd.Function):
    @staticmethod
    def forward(ctx, idx, *feats):
        M = len(feats)
        B, T, C = feats[0].shape
        out = torch.empty(B, T, M, C, dtype=feats[0].dtype, device=feats[0].device)
        L = _lib.lib()
        es = out.element_size()
        if B * T:
            with _on(out.device):
                for m in range(M):
                    check(L.r3d_exchange_one_fwd(_p(feats[m]), _p(feats[(m + 1) % M]), _p(idx[m]), idx.shape[1],
                                                 out.data_ptr() + m * C * es, M * C, B * T, C, _dt(out), _stream()))
        ctx.save_for_backward(idx)
        ctx.M = M
        return out

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        M = ctx.M
        g = g.contiguous()
        B, T, _, C = g.shape
        L = _lib.lib()
        es = g.element_size()
        outs = []
        with _on(g.device):
            for j in range(M):
                dx = torch.empty(B, T, C, dtype=g.dtype, device=g.device)
                if B * T:
                    p = (j - 1) % M
                    check(L.r3d_exchange_one_bwd(g.data_ptr() + j * C * es, g.data_ptr() + p * C * es, M * C, _p(idx[j]),
                                                 _p(idx[p]), idx.shape[1], _p(dx), B * T, C, _dt(g), _stream()))
                outs.append(dx)
        return (None, *outs)


def exchange_multi(feats, idx: torch.Tensor) -> torch.Tensor:
    """M tensors (B,T,C) + idx (M, k) int64 -> stacked (B,T,M,C): stream m takes the channels idx[m] from modality
    (m + 1) mod M.  For M = 2 this is ops.exchange with BLEND_SWAP."""
    feats = [f.contiguous() for f in feats]
    _need_cuda(*feats, idx)
    for f in feats[1:]:
        if f.shape != feats[0].shape or f.dtype != feats[0].dtype:
            raise R3DError("all modalities must agree in shape and dtype")
    C = feats[0].shape[-1]
    if C % (4 if feats[0].dtype == torch.float32 else 8):
        raise R3DError("exchange_multi needs C to be a multiple of the 128-bit vector width")
    return _ExchangeMulti.apply(idx.to(torch.int64).contiguous(), *feats)


class _MTokenAttn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, M, heads):
        rows = qkv.shape[0]
        C = qkv.shape[-1] // 3
        qkv = qkv.contiguous()
        out = torch.empty(rows, M, C, dtype=qkv.dtype, device=qkv.device)
        with _on(qkv.device):
            check(_lib.lib().r3d_mtoken_attn_fwd(_p(qkv), _p(out), rows, M, C, heads, _dt(qkv), _stream()))
        ctx.save_for_backward(qkv)
        ctx.meta = (M, heads)
        return out

    @staticmethod
    def backward(ctx, g):
        (qkv,) = ctx.saved_tensors
        M, heads = ctx.meta
        rows, C = qkv.shape[0], qkv.shape[-1] // 3
        dqkv = torch.empty_like(qkv)
        with _on(qkv.device):
            check(_lib.lib().r3d_mtoken_attn_bwd(_p(qkv), _p(g.contiguous()), _p(dqkv), rows, M, C, heads, _dt(qkv),
                                                 _stream()))
        return dqkv, None, None


def mtoken_attention(qkv: torch.Tensor, heads: int) -> torch.Tensor:
    """qkv (R, M, 3C) -> (R, M, C): per head, token m attends to the OTHER M - 1 tokens (diagonal masked with -inf,
    tokenfusion.py:68-72), scale head_dim^-0.5 (transformerblock.py:11,27)."""
    _need_cuda(qkv)
    _dt(qkv)
    return _MTokenAttn.apply(qkv, qkv.shape[1], int(heads))


class _TokenMean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        R, M, C = x.shape
        x = x.contiguous()
        out = torch.empty(R, C, dtype=x.dtype, device=x.device)
        with _on(x.device):
            check(_lib.lib().r3d_token_mean(_p(x), _p(out), R, M, C, _dt(x), 0, _stream()))
        ctx.M = M
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        R, C = g.shape
        dx = torch.empty(R, ctx.M, C, dtype=g.dtype, device=g.device)
        with _on(g.device):
            check(_lib.lib().r3d_token_mean(_p(g), _p(dx), R, ctx.M, C, _dt(g), 1, _stream()))
        return dx


def token_mean(x: torch.Tensor) -> torch.Tensor:
    """(R, M, C) -> (R, C): mean over the modality tokens (tokenfusion.py:95)."""
    _need_cuda(x)
    _dt(x)
    return _TokenMean.apply(x)


class _FusedBlockM(torch.autograd.Function):
    """The Block for M >= 3 modality tokens (a real (M-1)-way softmax per head): LN -> qkv GEMM -> M-token attention ->
    proj GEMM (+bias +residual) -> LN -> fc1 GEMM (GELU) -> fc2 GEMM (+bias +residual); hand-written kernels only."""

    @staticmethod
    def forward(ctx, x, n1w, n1b, qkv_w, qkv_b, wp, bp, n2w, n2b, w1, b1, w2, b2, eps1, eps2, heads):
        dt = x.dtype
        R, M, C = x.shape
        xc = x.contiguous().view(R * M, C)
        cast = lambda t: None if t is None else t.detach().to(dt).contiguous()
        n1w_, n1b_, n2w_, n2b_ = cast(n1w), cast(n1b), cast(n2w), cast(n2b)
        wq, bq, wp_, bp_, w1_, b1_, w2_, b2_ = cast(qkv_w), cast(qkv_b), cast(wp), cast(bp), cast(w1), cast(b1), cast(w2), cast(b2)
        L = _lib.lib()
        with _on(x.device):
            h1, m1, r1 = _ln_fwd_raw(xc, n1w_, n1b_, eps1, 0)
            qkv = gemm(h1, wq, bias=bq)                                   # (R*M, 3C)
            att = torch.empty(R * M, C, dtype=dt, device=x.device)
            check(L.r3d_mtoken_attn_fwd(_p(qkv), _p(att), R, M, C, heads, _dt(xc), _stream()))
            x1 = gemm(att, wp_, bias=bp_, residual=xc)
            h2, m2, r2 = _ln_fwd_raw(x1, n2w_, n2b_, eps2, 0)
            g1, H = gemm(h2, w1_, bias=b1_, act=ACT_GELU, want_aux=True)
            x2 = gemm(g1, w2_, bias=b2_, residual=x1)
        ctx.save_for_backward(xc, m1, r1, h1, qkv, att, x1, m2, r2, h2, H, g1, n1w_, wq, wp_, n2w_, w1_, w2_)
        ctx.meta = (x.shape, heads, qkv_b is not None,
                    [None if t is None else t.dtype for t in (n1w, n1b, qkv_w, qkv_b, wp, bp, n2w, n2b, w1, b1, w2, b2)])
        return x2.view(x.shape)

    @staticmethod
    def backward(ctx, dx2):
        xc, m1, r1, h1, qkv, att, x1, m2, r2, h2, H, g1, n1w, wq, wp, n2w, w1, w2 = ctx.saved_tensors
        xshape, heads, has_qkv_b, pdt = ctx.meta
        R, M, C = xshape
        dx2 = dx2.contiguous().view(-1, C)
        L = _lib.lib()
        with _on(dx2.device):
            db2 = colsum(dx2)
            dH, db1 = gemm(dx2, w2, True, False, aux_in=H, colsum=True)
            dW2 = gemm(dx2, g1, False, False)
            dh2 = gemm(dH, w1, True, False)
            dW1 = gemm(dH, h2, False, False)
            dx1, dg2, dbt2 = _ln_bwd_raw(dh2, x1, m2, r2, n2w, 0, dx2)
            dbp = colsum(dx1)
            datt = gemm(dx1, wp, True, False)
            dWp = gemm(dx1, att, False, False)
            dqkv = torch.empty_like(qkv)
            check(L.r3d_mtoken_attn_bwd(_p(qkv), _p(datt), _p(dqkv), R, M, C, heads, _dt(qkv), _stream()))
            dh1 = gemm(dqkv, wq, True, False)
            dWq = gemm(dqkv, h1, False, False)
            dbq = colsum(dqkv) if has_qkv_b else None
            dx, dg1, dbt1 = _ln_bwd_raw(dh1, xc, m1, r1, n1w, 0, dx1)
        g = [dg1, dbt1, dWq, dbq, dWp, dbp, dg2, dbt2, dW1, db1, dW2, db2]
        g = [None if (t is None or d is None) else t.to(d) for t, d in zip(g, pdt)]
        return (dx.view(xshape), *g, None, None, None)


def fused_block_multi(x, norm1, qkv, proj, norm2, fc1, fc2, heads):
    """x (R, M, C), M >= 3 -> (R, M, C), see _FusedBlockM."""
    return _FusedBlockM.apply(x, norm1.weight, norm1.bias, qkv.weight, qkv.bias, proj.weight, proj.bias, norm2.weight,
                              norm2.bias, fc1.weight, fc1.bias, fc2.weight, fc2.bias, norm1.eps, norm2.eps, int(heads))


# ---------------------------------------------------------------------------------
# N1: token-axis selection (north_star kernels 3-6; no reference symbol -- the reference ships the channel exchange
# only; oracle: oracle/fuser_oracle.py:token_fusion_tokens, parity unpinned)
# ---------------------------------------------------------------------------------
def token_scores(x: torch.Tensor, rtol: float = DEFAULT_RTOL, gram_impl: int = GRAM_TCGEN05, return_erank: bool = False):
    """(B, T, C) -> (B, T) float32: informativeness of every token, s_t = sum_j p_j u_{tj}^2 over the left singular
    vectors of each sample (p = sigma / sum sigma).  Runs the effective-rank chain and reuses what it saved."""
    _need_cuda(x)
    x = x.detach().contiguous()
    B, T, C = x.shape
    er, sigma, U, Y, _ = _erank_fwd_raw(x, rtol, gram_impl)
    out = torch.empty(B, T, dtype=torch.float32, device=x.device)
    with _on(x.device):
        check(_lib.lib().r3d_token_scores(_p(sigma), _p(U), _p(Y), B, T, C, rtol, _p(out), _stream()))
    return (out, er) if return_erank else out


def token_mask(idx_r: torch.Tensor, idx_d: torch.Tensor, T: int) -> torch.Tensor:
    """Per-sample index lists (B, k) int64 -> (B, T) uint8, bit 0 = token in S_rgb(b), bit 1 = token in S_depth(b)."""
    _need_cuda(idx_r, idx_d)
    B, k = idx_r.shape
    mask = torch.empty(B, T, dtype=torch.uint8, device=idx_r.device)
    with _on(idx_r.device):
        check(_lib.lib().r3d_token_mask(_p(idx_r.contiguous()), _p(idx_d.contiguous()), B, T, k, _p(mask), _stream()))
    return mask


class _TokenExchange(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rgb, depth, mask):
        B, T, C = rgb.shape
        out = torch.empty(B, T, 2, C, dtype=rgb.dtype, device=rgb.device)
        if B * T:
            with _on(rgb.device):
                check(_lib.lib().r3d_token_exchange_fwd(_p(rgb), _p(depth), _p(mask), _p(out), B * T, C, _dt(rgb),
                                                        _stream()))
        ctx.save_for_backward(mask)
        return out

    @staticmethod
    def backward(ctx, g):
        (mask,) = ctx.saved_tensors
        g = g.contiguous()
        B, T, _, C = g.shape
        d_rgb = torch.empty(B, T, C, dtype=g.dtype, device=g.device)
        d_dep = torch.empty(B, T, C, dtype=g.dtype, device=g.device)
        if B * T:
            with _on(g.device):
                check(_lib.lib().r3d_token_exchange_bwd(_p(g), _p(mask), _p(d_rgb), _p(d_dep), B * T, C, _dt(g),
                                                        _stream()))
        return d_rgb, d_dep, None


def token_exchange(rgb, depth, idx_r, idx_d) -> torch.Tensor:
    """(B,T,C) x2 + per-sample token index sets (B, k) -> stacked (B,T,2,C): tokens in S_rgb(b) of the rgb stream are
    replaced by the depth tokens at the same positions and vice versa; differentiable in rgb and depth."""
    _need_cuda(rgb, depth, idx_r, idx_d)
    rgb, depth = _btc(rgb, depth)
    _dt(rgb)
    B, T, C = rgb.shape
    if idx_r.shape != idx_d.shape or idx_r.dim() != 2 or idx_r.shape[0] != B:
        raise R3DError(f"token index sets must both be (B, k), got {tuple(idx_r.shape)} and {tuple(idx_d.shape)}")
    mask = token_mask(idx_r.to(torch.int64), idx_d.to(torch.int64), T)
    return _TokenExchange.apply(rgb, depth, mask)


def token_fusion_host(rgb: torch.Tensor, depth: torch.Tensor, k: int):
    """Host-buffer entry (pinned CPU tensors in, pinned CPU tensors out) through the
    C ABI's r3d_token_fusion_host: the `e2e` measurement path."""
    if rgb.is_cuda or depth.is_cuda:
        raise R3DError("token_fusion_host takes host tensors")
    B, T, C = rgb.shape
    out = torch.empty(B, T, 2, C, dtype=rgb.dtype).pin_memory()
    idx = torch.empty(2, max(k, 1), dtype=torch.int64).pin_memory()
    check(_lib.lib().r3d_token_fusion_host(_p(rgb), _p(depth), B, T, C, _dt(rgb), k, _p(out), _p(idx[0]), _p(idx[1]),
                                           _stream()))
    return out, idx[:, :k]


def fuser_step_host(rgb: torch.Tensor, depth: torch.Tensor, gst: Optional[torch.Tensor] = None, k: Optional[int] = None,
                    rtol: float = DEFAULT_RTOL, erank_weight: float = 1.0, out=None):
    """The whole hot path on HOST tensors through the single C entry r3d_fuser_step_host (H2D, erank of both
    modalities, score -> bottom-k -> exchange, backward of `gst` + the erank gradient, D2H).  Returns
    (stacked, erank (2B,), d_rgb, d_depth, idx (2, k)); gradients are None without `gst`.  `out`: optional tuple of
    preallocated (pinned) result tensors in that order."""
    if rgb.is_cuda or depth.is_cuda or (gst is not None and gst.is_cuda):
        raise R3DError("fuser_step_host takes host tensors")
    B, T, C = rgb.shape
    k = C // 4 if k is None else k
    if out is None:
        stacked = torch.empty(B, T, 2, C, dtype=rgb.dtype)
        er = torch.empty(2 * B, dtype=torch.float32)
        d_r = torch.empty_like(rgb) if gst is not None else None
        d_d = torch.empty_like(rgb) if gst is not None else None
        idx = torch.empty(2, max(k, 1), dtype=torch.int64)
    else:
        stacked, er, d_r, d_d, idx = out
    check(_lib.lib().r3d_fuser_step_host(_p(rgb), _p(depth), _p(gst), B, T, C, _dt(rgb), k, float(rtol), float(erank_weight),
                                         _p(stacked), _p(er), _p(d_r), _p(d_d), _p(idx), _stream()))
    return stacked, er, d_r, d_d, idx[:, :k]


# ---------------------------------------------------------------------------------
# torch.ops.r3d.* registration.  Thin wrappers over the functions above.
# ---------------------------------------------------------------------------------
def _register():
    try:
        from torch.library import custom_op
    except Exception:  # pragma: no cover
        return

    @custom_op("r3d::channel_score", mutates_args=(), device_types="cuda")
    def _cs(rgb: torch.Tensor, depth: torch.Tensor) -> torch.Tensor:
        return channel_score(rgb, depth)

    @_cs.register_fake
    def _(rgb, depth):
        return rgb.new_empty((2, rgb.shape[-1]), dtype=torch.float32)

    @custom_op("r3d::bottomk", mutates_args=(), device_types="cuda")
    def _bk(score: torch.Tensor, k: int) -> torch.Tensor:
        return bottomk(score, k)

    @_bk.register_fake
    def _(score, k):
        return score.new_empty((*score.shape[:-1], k), dtype=torch.int64)

    @custom_op("r3d::exchange_fwd", mutates_args=(), device_types="cuda")
    def _ef(rgb: torch.Tensor, depth: torch.Tensor, idx_r: torch.Tensor, idx_d: torch.Tensor,
            alpha: Optional[torch.Tensor], blend_mode: int) -> torch.Tensor:
        a = None if alpha is None else alpha.reshape(-1).float().contiguous()
        return _exchange_fwd_raw(rgb.contiguous(), depth.contiguous(), idx_r.contiguous(), idx_d.contiguous(), a,
                                 None, blend_mode)

    @_ef.register_fake
    def _(rgb, depth, idx_r, idx_d, alpha, blend_mode):
        B, T, C = rgb.shape
        return rgb.new_empty((B, T, 2, C))

    @custom_op("r3d::erank", mutates_args=(), device_types="cuda")
    def _er(x: torch.Tensor, rtol: float) -> torch.Tensor:
        return _erank_fwd_raw(x.contiguous(), rtol, GRAM_TCGEN05)[0]

    @_er.register_fake
    def _(x, rtol):
        return x.new_empty((x.shape[0],), dtype=torch.float32)

    # ---- the remaining schemas of SURVEY.md 8(b): exchange_bwd, erank_fwd (tuple), erank_bwd, with autograd wired
    # through torch.library (register_autograd), so the ops are differentiable when called as torch.ops.r3d.*
    @custom_op("r3d::exchange_bwd", mutates_args=(), device_types="cuda")
    def _eb(g: torch.Tensor, rgb: torch.Tensor, depth: torch.Tensor, idx_r: torch.Tensor, idx_d: torch.Tensor,
            alpha: Optional[torch.Tensor], blend_mode: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        a = None if alpha is None else alpha.reshape(-1).float().contiguous()
        swap = blend_mode == BLEND_SWAP
        d_r, d_d, cs = _exchange_bwd_raw(g.contiguous(), None if swap else rgb.contiguous(),
                                         None if swap else depth.contiguous(), idx_r.contiguous(), idx_d.contiguous(), a,
                                         None, None, blend_mode)
        d_alpha = torch.zeros(rgb.shape[-1], dtype=torch.float32, device=g.device) if cs is None else cs[0].clone()
        return d_r, d_d, d_alpha

    @_eb.register_fake
    def _(g, rgb, depth, idx_r, idx_d, alpha, blend_mode):
        B, T, _, C = g.shape
        return g.new_empty((B, T, C)), g.new_empty((B, T, C)), g.new_empty((C,), dtype=torch.float32)

    def _ef_setup(ctx, inputs, output):
        rgb, depth, idx_r, idx_d, alpha, blend_mode = inputs
        ctx.save_for_backward(rgb, depth, idx_r, idx_d, alpha if alpha is not None else rgb.new_empty(0))
        ctx.has_alpha, ctx.blend = alpha is not None, blend_mode

    def _ef_backward(ctx, g):
        rgb, depth, idx_r, idx_d, alpha = ctx.saved_tensors
        d_r, d_d, d_a = torch.ops.r3d.exchange_bwd(g, rgb, depth, idx_r, idx_d, alpha if ctx.has_alpha else None, ctx.blend)
        return d_r, d_d, None, None, (d_a.reshape(alpha.shape).to(alpha.dtype) if ctx.has_alpha else None), None

    _ef.register_autograd(_ef_backward, setup_context=_ef_setup)

    @custom_op("r3d::erank_fwd", mutates_args=(), device_types="cuda")
    def _erf(x: torch.Tensor, rtol: float) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        er, sigma, U, Y, _ = _erank_fwd_raw(x.contiguous(), rtol, GRAM_TCGEN05)
        return er, sigma, U, Y

    @_erf.register_fake
    def _(x, rtol):
        B, T, C = x.shape
        n, m = min(T, C), max(T, C)
        f = dict(dtype=torch.float32)
        return x.new_empty((B,), **f), x.new_empty((B, n), **f), x.new_empty((B, n, n), **f), x.new_empty((B, n, m), **f)

    @custom_op("r3d::erank_bwd", mutates_args=(), device_types="cuda")
    def _erb(g: torch.Tensor, x: torch.Tensor, sigma: torch.Tensor, U: torch.Tensor, Y: torch.Tensor, erank: torch.Tensor,
             rtol: float) -> torch.Tensor:
        B, T, C = x.shape
        L = _lib.lib()
        dt = _dt(x)
        with _on(x.device):
            ws = torch.empty(L.r3d_erank_workspace_bytes(B, T, C, dt), dtype=torch.uint8, device=x.device)
            dx = torch.empty_like(x)
            check(L.r3d_erank_bwd(_p(g.contiguous().float()), _p(erank), _p(sigma), _p(U), _p(Y), B, T, C, dt, rtol, _p(ws),
                                  _p(dx), 0, _stream()))
        return dx

    @_erb.register_fake
    def _(g, x, sigma, U, Y, erank, rtol):
        return torch.empty_like(x)

    def _erf_setup(ctx, inputs, output):
        x, rtol = inputs
        er, sigma, U, Y = output
        ctx.save_for_backward(x, sigma, U, Y, er)
        ctx.rtol = rtol

    def _erf_backward(ctx, g_er, g_sigma, g_U, g_Y):
        x, sigma, U, Y, er = ctx.saved_tensors
        return torch.ops.r3d.erank_bwd(g_er, x, sigma, U, Y, er, ctx.rtol), None

    _erf.register_autograd(_erf_backward, setup_context=_erf_setup)


_register()


# ---------------------------------------------------------------------------------
# One whole pass of the hot path over a batch, preallocated (what bench.py times):
#   forward : erank(rgb), erank(depth) ; score -> [all-reduce] -> bottom-k -> exchange
#   backward: exchange backward of an upstream gradient + d(mean erank)/dX, accumulated
# ---------------------------------------------------------------------------------
class FuserStep:
    """Preallocated buffers + the call sequence for one fused fwd/bwd step on (2, B, T, C)
    inputs (rgb = buf[0], depth = buf[1]).  Tokenfusion variant, eval-branch score."""

    def __init__(self, B, T, C, dtype, device, k=None, rtol=DEFAULT_RTOL, gram_impl=GRAM_TCGEN05, group=None,
                 erank_weight=1.0):
        self.B, self.T, self.C, self.dtype, self.device = B, T, C, dtype, device
        self.k = C // 4 if k is None else k
        self.rtol, self.gram_impl, self.group, self.w = float(rtol), int(gram_impl), group, float(erank_weight)
        L = _lib.lib()
        rows = B * T
        n, m = (T, C) if T < C else (C, T)
        f32 = dict(dtype=torch.float32, device=device)
        dt = _DT[dtype]
        self.ws_score = torch.empty(L.r3d_score_workspace_floats(rows, C), **f32)
        self.packed = torch.zeros(2 * C + 2, **f32)          # [sum|rgb| (C) | sum|depth| (C) | sum erank | count]
        self.score = torch.empty(2, C, **f32)
        self.idx = torch.empty(2, max(self.k, 1), dtype=torch.int64, device=device)
        self.out = torch.empty(B, T, 2, C, dtype=dtype, device=device)
        self.dgrad = torch.empty(2, B, T, C, dtype=dtype, device=device)
        self.ws_er = torch.empty(L.r3d_erank_workspace_bytes(2 * B, T, C, dt), dtype=torch.uint8, device=device)
        self.er = torch.empty(2 * B, **f32)
        self.sigma = torch.empty(2 * B, n, **f32)
        self.U = torch.empty(2 * B, n, n, **f32)
        self.Y = torch.empty(2 * B, n, m, **f32)
        self.sweeps = torch.empty(2 * B, dtype=torch.int32, device=device)
        self.gvec = torch.empty(2 * B, **f32)
        self.world = 1
        if group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(group)
        self.gvec.fill_(self.w / float(2 * B * self.world))   # d(mean erank over the global batch)/d erank_b

    def __call__(self, buf: torch.Tensor, gst: torch.Tensor):
        B, T, C = self.B, self.T, self.C
        L = _lib.lib()
        dt = _DT[self.dtype]
        st = _stream()
        rows = B * T
        rgb_p, dep_p = buf.data_ptr(), buf.data_ptr() + buf.stride(0) * buf.element_size()
        # ---- forward: effective rank of both modalities (2B samples in one batch)
        check(L.r3d_erank_fwd(buf.data_ptr(), 2 * B, T, C, dt, self.rtol, self.gram_impl, _p(self.ws_er), _p(self.er),
                              _p(self.sigma), _p(self.U), _p(self.Y), _p(self.sweeps), st))
        # ---- forward: score sums -> (all-reduce with the erank statistic) -> bottom-k -> exchange
        check(L.r3d_channel_score_partial(rgb_p, dep_p, rows, C, dt, _p(self.ws_score), st))
        # packed = [sum|rgb| | sum|depth| | sum erank | rows], all written by the finalize kernel (no ATen kernels, no
        # host copy: graph-capturable); the quotient sums / rows is formed inside the bottom-k kernel
        check(L.r3d_score_finalize_packed(_p(self.ws_score), rows, C, _p(self.er), 2 * B, _p(self.packed), st))
        if self.world > 1:
            torch.distributed.all_reduce(self.packed, group=self.group)
        check(L.r3d_bottomk_scaled(_p(self.packed), 2, C, self.k, self.packed.data_ptr() + (2 * C + 1) * 4,
                                   _p(self.idx), _p(self.score), st))
        ir, idd = self.idx[0].data_ptr(), self.idx[1].data_ptr()
        check(L.r3d_exchange_fwd(rgb_p, dep_p, ir, idd, self.k, None, None, BLEND_SWAP, _p(self.out), rows, C, dt, st))
        # ---- backward: exchange backward into dgrad, then accumulate d(mean erank)/dX
        d0, d1 = self.dgrad.data_ptr(), self.dgrad.data_ptr() + self.dgrad.stride(0) * self.dgrad.element_size()
        check(L.r3d_exchange_bwd(_p(gst), None, None, ir, idd, self.k, None, None, None, BLEND_SWAP, d0, d1, None,
                                 rows, C, dt, st))
        check(L.r3d_erank_bwd(_p(self.gvec), _p(self.er), _p(self.sigma), _p(self.U), _p(self.Y), 2 * B, T, C, dt,
                              self.rtol, _p(self.ws_er), _p(self.dgrad), 1, st))
        return self.out, self.er, self.dgrad



class FuserTrainStep:
    """One TRAINING step of the fuser path on (2, B, T, C) inputs (rgb = buf[0], depth = buf[1]) -- what bench.py times:

      forward : effective rank of both modalities (2B samples, the collapse statistic / regulariser);
                CMFuser.forward = channel score -> [one all-reduce of the packed statistic in global scope] -> bottom-k
                -> exchange -> Block (LayerNorm, V / proj / MLP GEMMs) -> LayerNorm -> mean over the two tokens;
      backward: an upstream gradient through CMFuser (input and parameter gradients), d(mean erank)/dX accumulated
                into the input gradients, and -- with more than one rank -- the all-reduce of the fuser's parameter
                gradients (dist.GradBucket; the reference's nn.DataParallel reduce-add, main_utkinects.py:129) on a side
                stream, overlapped with the effective-rank backward.
    """

    def __init__(self, fuser, B, T, C, dtype, device, rtol=DEFAULT_RTOL, gram_impl=GRAM_TCGEN05, group=None,
                 erank_weight=1.0):
        from . import dist as D
        self.fuser, self.B, self.T, self.C, self.dtype, self.device = fuser, B, T, C, dtype, device
        self.rtol, self.gram_impl, self.group, self.w = float(rtol), int(gram_impl), group, float(erank_weight)
        L = _lib.lib()
        n, m = (T, C) if T < C else (C, T)
        f32 = dict(dtype=torch.float32, device=device)
        dt = _DT[dtype]
        self.ws_er = torch.empty(L.r3d_erank_workspace_bytes(2 * B, T, C, dt), dtype=torch.uint8, device=device)
        self.er = torch.empty(2 * B, **f32)
        self.sigma = torch.empty(2 * B, n, **f32)
        self.U = torch.empty(2 * B, n, n, **f32)
        self.Y = torch.empty(2 * B, n, m, **f32)
        self.sweeps = torch.empty(2 * B, dtype=torch.int32, device=device)
        self.world = 1
        if group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(group)
        self.gvec = torch.full((2 * B,), self.w / float(2 * B * self.world), **f32)   # d(global mean erank)/d erank_b
        self.bucket = D.GradBucket(fuser.parameters(), group=group) if self.world > 1 else None
        self.side = torch.cuda.Stream(device=device) if self.world > 1 else None
        self.allreduce_ms = None

    def __call__(self, buf: torch.Tensor, gy: torch.Tensor):
        B, T, C = self.B, self.T, self.C
        L = _lib.lib()
        dt = _DT[self.dtype]
        st = _stream()
        check(L.r3d_erank_fwd(buf.data_ptr(), 2 * B, T, C, dt, self.rtol, self.gram_impl, _p(self.ws_er), _p(self.er),
                              _p(self.sigma), _p(self.U), _p(self.Y), _p(self.sweeps), st))
        r = buf[0].detach().requires_grad_(True)
        d = buf[1].detach().requires_grad_(True)
        for p in self.fuser.parameters():             # what optimizer.zero_grad(set_to_none=True) does in a training loop
            p.grad = None
        self.fuser.statistic_extra = self.er          # its sum rides in the packed score buffer (one all-reduce)
        y = self.fuser({"rgb": r, "depth": d}, "test")
        y.backward(gy)
        main = torch.cuda.current_stream()
        if self.bucket is not None:                  # gradient all-reduce beside the effective-rank backward
            self.side.wait_stream(main)
            with torch.cuda.stream(self.side):
                self.bucket.allreduce(average=True)
        n, m = (T, C) if T < C else (C, T)
        for h, g in ((0, r.grad), (1, d.grad)):      # per modality: B samples each, accumulated into the input gradient
            o = h * B
            check(L.r3d_erank_bwd(self.gvec.data_ptr() + o * 4, self.er.data_ptr() + o * 4,
                                  self.sigma.data_ptr() + o * n * 4, self.U.data_ptr() + o * n * n * 4,
                                  self.Y.data_ptr() + o * n * m * 4, B, T, C, dt, self.rtol, _p(self.ws_er), _p(g), 1, st))
        if self.bucket is not None:
            main.wait_stream(self.side)
        return y, self.er, r.grad, d.grad
