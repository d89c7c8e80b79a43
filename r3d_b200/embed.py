"""The callers on the input side of the fuser on hand-written kernels:

* ``RGBEmbed``   -- ``FUTR.input_embed`` + ``F.relu``  (model/futr_safuser_tokenfusion.py:111,179,183)
* ``DepthEmbed`` -- ``FUTR.depth_projection`` + ``depth_layernorm`` + ``F.relu``  (tokenfusion.py:143,147,194-197)

Both run the tcgen05 GEMM of csrc/linear_tcgen05.cu and emit, as a by-product, the column sums of |output| -- the
channel score of tokenfusion.py:49-50 -- so that ``CMFuser.forward(..., score_parts=...)`` can skip its own score pass
(a1): the rgb sums come out of the GEMM epilogue, the depth sums out of the LayerNorm+ReLU kernel.

Parameter names equal the reference's (``input_embed.*``, ``depth_projection.*``, ``depth_layernorm.*``), so a slice of
a reference ``FUTR`` state_dict loads with ``strict=False``; initialisation follows tokenfusion.py:116,145
(xavier-uniform weights).  ``FuserFront`` bundles the two with a ``CMFuser`` the way ``FUTR.forward`` wires them
(tokenfusion.py:179-199).
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch
from torch import nn

from . import _lib, ops
from ._lib import R3DError, check
from .ops import _dt, _on, _p, _stream


class ScoreParts:
    """Column-sum partials of |x| over the rows of one modality: (parts, C) float32, rows counted."""
    __slots__ = ("partial", "parts", "rows")

    def __init__(self, partial: torch.Tensor, parts: int, rows: int):
        self.partial, self.parts, self.rows = partial, int(parts), int(rows)

    def sums(self) -> torch.Tensor:
        C = self.partial.shape[-1]
        out = torch.empty(C, dtype=torch.float32, device=self.partial.device)
        with _on(out.device):
            check(_lib.lib().r3d_colsum_finalize(_p(self.partial), self.parts, C, 0, _p(out), _stream()))
        return out


def pack_score_parts(rgb: ScoreParts, depth: ScoreParts, er: Optional[torch.Tensor] = None) -> torch.Tensor:
    """-> (2C + 2,) packed statistic [sum|rgb| | sum|depth| | sum er | rows] (ops.bottomk_packed consumes it)."""
    if rgb.rows != depth.rows:
        raise R3DError("rgb and depth score partials cover different row counts")
    C = rgb.partial.shape[-1]
    packed = torch.empty(2 * C + 2, dtype=torch.float32, device=rgb.partial.device)
    e = None if er is None else er.reshape(-1).float().contiguous()
    with _on(packed.device):
        check(_lib.lib().r3d_score_pack(_p(rgb.partial), rgb.parts, _p(depth.partial), depth.parts, rgb.rows, C, _p(e),
                                        0 if e is None else e.numel(), _p(packed), _stream()))
    return packed


class _LinearReLU(torch.autograd.Function):
    """y = relu(x W^T + b) with the |y| column-sum partials from the GEMM epilogue (f3)."""

    @staticmethod
    def forward(ctx, x2d, weight, bias):
        dt = x2d.dtype
        w = weight.detach().to(dt).contiguous()
        b = None if bias is None else bias.detach().to(dt).contiguous()
        R, K = x2d.shape
        N = w.shape[0]
        L = _lib.lib()
        dev = x2d.device
        with _on(dev):
            y = torch.empty(R, N, dtype=dt, device=dev)
            parts = (R + 127) // 128
            cs = torch.empty(parts, N, dtype=torch.float32, device=dev)
            ws = torch.empty(L.r3d_gemm_workspace_bytes(R, N, K, _dt(x2d)), dtype=torch.uint8, device=dev)
            ep = _lib.Epilogue()
            ep.bias = None if b is None else b.data_ptr()
            ep.colsum_partial = cs.data_ptr()
            ep.act, ep.colsum_abs = ops.ACT_RELU, 1
            check(L.r3d_gemm(_p(x2d), _p(w), _p(y), R, N, K, 1, 1, _dt(x2d), ctypes.byref(ep), _p(ws), _stream()))
        ctx.save_for_backward(x2d, w, y)
        ctx.meta = (weight.dtype, None if bias is None else bias.dtype, x2d.requires_grad)
        ctx.mark_non_differentiable(cs)
        return y, cs

    @staticmethod
    def backward(ctx, dy, _dcs):
        x2d, w, y = ctx.saved_tensors
        wdt, bdt, need_dx = ctx.meta
        dy = dy.contiguous()
        R, N = y.shape
        L = _lib.lib()
        with _on(dy.device):
            ws = torch.empty(L.r3d_colsum_workspace_floats(R, N), dtype=torch.float32, device=dy.device)
            dpre = torch.empty_like(y)
            db = torch.empty(N, dtype=y.dtype, device=dy.device)
            check(L.r3d_relu_bwd(_p(dy), _p(y), R, N, _dt(y), _p(ws), _p(dpre), _p(db), _stream()))
            dW = ops.gemm(dpre, x2d, False, False)
            dx = ops.gemm(dpre, w, True, False) if need_dx else None
        return dx, dW.to(wdt), None if bdt is None else db.to(bdt)


class _LinearLNReLU(torch.autograd.Function):
    """y = relu(LayerNorm(x W^T + b)) with the |y| column-sum partials from the LayerNorm kernel (f2)."""

    @staticmethod
    def forward(ctx, x2d, weight, bias, gamma, beta, eps):
        dt = x2d.dtype
        w = weight.detach().to(dt).contiguous()
        b = None if bias is None else bias.detach().to(dt).contiguous()
        g_, b_ = gamma.detach().to(dt).contiguous(), beta.detach().to(dt).contiguous()
        R, K = x2d.shape
        N = w.shape[0]
        L = _lib.lib()
        dev = x2d.device
        with _on(dev):
            pre = ops.gemm(x2d, w, bias=b)                               # (R, N), the projection
            y = torch.empty_like(pre)
            mean = torch.empty(R, dtype=torch.float32, device=dev)
            rstd = torch.empty(R, dtype=torch.float32, device=dev)
            parts = int(L.r3d_ln_relu_parts(R))
            cs = torch.empty(parts, N, dtype=torch.float32, device=dev)
            check(L.r3d_ln_relu_fwd(_p(pre), _p(g_), _p(b_), R, N, _dt(pre), float(eps), _p(y), _p(mean), _p(rstd), _p(cs),
                                    _stream()))
        ctx.save_for_backward(x2d, w, pre, mean, rstd, g_, b_)
        ctx.meta = (weight.dtype, None if bias is None else bias.dtype, gamma.dtype, beta.dtype, x2d.requires_grad)
        ctx.mark_non_differentiable(cs)
        return y, cs

    @staticmethod
    def backward(ctx, dy, _dcs):
        x2d, w, pre, mean, rstd, g_, b_ = ctx.saved_tensors
        wdt, bdt, gdt, btdt, need_dx = ctx.meta
        dy = dy.contiguous()
        R, N = pre.shape
        L = _lib.lib()
        with _on(dy.device):
            dpre = torch.empty_like(pre)
            ws = torch.empty(L.r3d_ln_bwd_workspace_floats(R, N), dtype=torch.float32, device=dy.device)
            dgb = torch.empty(2, N, dtype=torch.float32, device=dy.device)
            check(L.r3d_ln_relu_bwd(_p(dy), _p(pre), _p(mean), _p(rstd), _p(g_), _p(b_), R, N, _dt(pre), _p(dpre), _p(ws),
                                    _p(dgb), _stream()))
            db = ops.colsum(dpre) if bdt is not None else None
            dW = ops.gemm(dpre, x2d, False, False)
            dx = ops.gemm(dpre, w, True, False) if need_dx else None
        return (dx, dW.to(wdt), None if db is None else db.to(bdt), dgb[0].to(gdt), dgb[1].to(btdt), None)


def _supported(x2d: torch.Tensor, N: int) -> bool:
    return x2d.is_cuda and x2d.dtype in (torch.float32, torch.bfloat16) and x2d.shape[1] % 8 == 0 and N % 8 == 0 \
        and x2d.data_ptr() % 16 == 0


class RGBEmbed(nn.Module):
    """``relu(input_embed(features))`` -- tokenfusion.py:111,179,183.  forward -> (B, S, C); ``last_score`` holds the
    |output| column-sum partials of the call (ScoreParts)."""

    def __init__(self, input_dim: int, hidden_dim: int):
        super().__init__()
        self.input_embed = nn.Linear(input_dim, hidden_dim)
        nn.init.xavier_uniform_(self.input_embed.weight)                 # tokenfusion.py:116
        self.last_score: Optional[ScoreParts] = None

    def forward(self, features: torch.Tensor) -> torch.Tensor:
        B, S, K = features.shape
        N = self.input_embed.out_features
        x2d = features.reshape(B * S, K)
        if not _supported(x2d, N):
            raise R3DError("RGBEmbed runs on CUDA fp32 / bf16 tensors with feature sizes that are multiples of 8 "
                           "(no CPU fallback)")
        y, cs = _LinearReLU.apply(x2d.contiguous(), self.input_embed.weight, self.input_embed.bias)
        self.last_score = ScoreParts(cs, cs.shape[0], B * S)
        return y.view(B, S, N)


class DepthEmbed(nn.Module):
    """``relu(depth_layernorm(depth_projection(depth.view(B, S, -1))))`` -- tokenfusion.py:143,147,194-197."""

    def __init__(self, in_features: int, hidden_dim: int):
        super().__init__()
        self.depth_projection = nn.Linear(in_features, hidden_dim)       # 224 * 224 (or 160 * 120) -> C
        nn.init.xavier_uniform_(self.depth_projection.weight)            # tokenfusion.py:146
        self.depth_layernorm = nn.LayerNorm(hidden_dim)
        self.last_score: Optional[ScoreParts] = None

    def forward(self, depth: torch.Tensor) -> torch.Tensor:
        B, S = depth.shape[:2]
        x2d = depth.reshape(B * S, -1)
        N = self.depth_projection.out_features
        if not _supported(x2d, N) or not ops.layer_norm_supported(x2d.new_empty(1, N), N):
            raise R3DError("DepthEmbed runs on CUDA fp32 / bf16 tensors with sizes the kernels cover (no CPU fallback)")
        ln = self.depth_layernorm
        y, cs = _LinearLNReLU.apply(x2d.contiguous(), self.depth_projection.weight, self.depth_projection.bias, ln.weight,
                                    ln.bias, ln.eps)
        self.last_score = ScoreParts(cs, cs.shape[0], B * S)
        return y.view(B, S, N)


class FuserFront(nn.Module):
    """RGB embedding + depth projection + CMFuser wired as in ``FUTR.forward`` (tokenfusion.py:179-199); the channel
    score reaches the fuser through the two producers' by-products, so no separate score pass runs."""

    def __init__(self, input_dim: int, depth_features: int, hidden_dim: int, n_head: int = 8, **fuser_kw):
        super().__init__()
        from .fuser import CMFuser
        self.rgb = RGBEmbed(input_dim, hidden_dim)
        self.depth = DepthEmbed(depth_features, hidden_dim)
        self.fuser = CMFuser(dim=hidden_dim, depth=1, num_heads=n_head, **fuser_kw)

    def forward(self, features: torch.Tensor, depth: torch.Tensor, mode: str = "test") -> torch.Tensor:
        src = self.rgb(features)
        dep = self.depth(depth)
        return self.fuser({"rgb": src, "depth": dep}, mode, score_parts=(self.rgb.last_score, self.depth.last_score))
