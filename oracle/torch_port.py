"""torch-CPU port of the reference fuser -- TEST INFRASTRUCTURE ONLY.

Purpose: (1) the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``
time this on the GPU box's host cores (the reference is Python and cannot
travel to the box, so the port stands in for it: ``kind: "port"``);
(2) the tests use its autograd as the gradient oracle for the wrapper.

It restates, op for op, what the reference modules execute -- including the
work the CUDA path proves dead (full qkv GEMM, 2x2 softmax with -inf diagonal,
clone + index_put + stack) -- so the CPU number is the reference's cost, not
the cost of a smarter algorithm.  Pinned against ``tests/golden/*.npz`` (made
from the unmodified reference) by ``tests/test_oracle_golden.py``.

Citations are relative to /root/reference.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn


class PortAttention(nn.Module):
    """model/extras/transformerblock.py:7-36."""

    def __init__(self, dim, num_heads, qkv_bias=False):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x, attn_mask):
        R, N, C = x.shape
        qkv = self.qkv(x).reshape(R, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        w = (q @ k.transpose(-2, -1)) * self.scale + attn_mask
        w = w.softmax(dim=-1)
        return self.proj((w @ v).transpose(1, 2).reshape(R, N, C)), w


class PortMLP(nn.Module):
    """model/extras/transformerblock.py:79-93 (named ``mlp.mlp.{0,2}`` like the reference)."""

    def __init__(self, dim, hidden):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(dim, hidden), nn.GELU(), nn.Linear(hidden, dim), nn.Dropout(0.0))

    def forward(self, x):
        return self.mlp(x)


class PortBlock(nn.Module):
    """model/extras/transformerblock.py:118-135."""

    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn = PortAttention(dim, num_heads, qkv_bias)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = PortMLP(dim, int(dim * mlp_ratio))

    def forward(self, x, attn_mask):
        a, w = self.attn(self.norm1(x), attn_mask)
        x = x + a
        return x + self.mlp(self.norm2(x)), w


class PortCMFuser(nn.Module):
    """The four reference fusers behind one class.

    variant 'tokenfusion': model/futr_safuser_tokenfusion.py:17-97
    variant 'vary'       : model/futr_safuser_tokenfusion_vary.py:17-87
    variant 'batchnorm'  : model/futr_safuser_batchnormalization.py:17-107
    variant 'safuser'    : model/futr_safuser_depth.py:17-64
    state_dict names equal the reference's, so reference weights load directly.
    """

    def __init__(self, dim, depth=1, num_heads=4, mlp_ratio=4.0, qkv_bias=False, variant="tokenfusion",
                 tie_break="index"):
        super().__init__()
        self.variant = variant
        self.tie_break = tie_break
        self.blocks = nn.ModuleList([PortBlock(dim, num_heads, mlp_ratio, qkv_bias) for _ in range(depth)])
        self.norm = nn.LayerNorm(dim)
        self.embd_drop = nn.Dropout(0.1)
        self.modality_token = nn.Parameter(torch.randn(1, 1, 1, dim))
        self.projection = nn.Linear(dim, dim)
        if variant != "safuser":
            self.fusion_conv = nn.Conv2d(2, 1, kernel_size=1)
        if variant == "vary":
            self.alpha = nn.Parameter(torch.ones(1, 1, dim))
        if variant == "batchnorm":
            self.alpha = nn.Parameter(torch.rand(1, 1, dim))
            self.bn_rgb = nn.BatchNorm1d(dim, affine=True)
            self.bn_depth = nn.BatchNorm1d(dim, affine=True)

    def _bottomk(self, score, k):
        if self.tie_break == "torch":           # exactly the reference call (tokenfusion.py:53)
            return torch.topk(score, k, dim=-1, largest=False)[1].reshape(-1)
        return torch.sort(score.reshape(-1), stable=True)[1][:k]   # ties -> lower index

    def token_fusion(self, rgb, depth, mode):
        B, T, C = rgb.shape
        if self.variant == "tokenfusion":
            if mode == "train":
                # tokenfusion.py:40-45 -- including the wasted partial backward
                loss = rgb.mean() + depth.mean()
                if loss.requires_grad:
                    g = torch.autograd.grad(loss, [rgb, depth], retain_graph=True)
                    s_r = g[0].abs().mean(dim=(0, 1))
                    s_d = g[1].abs().mean(dim=(0, 1))
                else:
                    s_r = s_d = torch.full((C,), 1.0 / (B * T * C), device=rgb.device)
            else:
                s_r, s_d = rgb.abs().mean(dim=(0, 1)), depth.abs().mean(dim=(0, 1))
            k = C // 4
        elif self.variant == "vary":
            s_r, s_d = rgb.abs().mean(dim=(0, 1)), depth.abs().mean(dim=(0, 1))
            k = C // 4
        else:
            rgb = self.bn_rgb(rgb.permute(0, 2, 1)).permute(0, 2, 1)
            depth = self.bn_depth(depth.permute(0, 2, 1)).permute(0, 2, 1)
            s_r, s_d = self.bn_rgb.weight.abs(), self.bn_depth.weight.abs()
            k = max(0, int(C * 0.1))
        i_r, i_d = self._bottomk(s_r.detach(), k), self._bottomk(s_d.detach(), k)
        ex_r, ex_d = rgb.clone(), depth.clone()
        if self.variant == "tokenfusion":
            ex_r[:, :, i_r] = depth[:, :, i_r]
            ex_d[:, :, i_d] = rgb[:, :, i_d]
        elif self.variant == "vary":
            ex_r[:, :, i_r] = self.alpha[:, :, i_r] * depth[:, :, i_r]
            ex_d[:, :, i_d] = self.alpha[:, :, i_d] * rgb[:, :, i_d]
        else:
            a_r, a_d = self.alpha[:, :, i_r], self.alpha[:, :, i_d]
            ex_r[:, :, i_r] = a_r * rgb[:, :, i_r] + (1 - a_r) * depth[:, :, i_r]
            ex_d[:, :, i_d] = a_d * depth[:, :, i_d] + (1 - a_d) * rgb[:, :, i_d]
        self.last_indices = (i_r, i_d)
        return torch.stack([ex_r, ex_d], dim=2)

    def forward(self, modal_feats, mode="test"):
        rgb, depth = modal_feats["rgb"], modal_feats["depth"]
        B, T, C = rgb.shape
        mask = torch.zeros(2, 2, device=rgb.device, dtype=rgb.dtype).masked_fill(
            torch.eye(2, device=rgb.device) == 1, float("-inf"))
        if self.variant == "safuser":
            st = torch.stack([rgb, depth], dim=2) + self.modality_token
        else:
            st = self.token_fusion(rgb, depth, mode)
        x = self.embd_drop(st.reshape(B * T, 2, C))
        x_res = x
        attns = []
        for blk in self.blocks:
            x, w = blk(x, mask)
            attns.append(w.view(B, T, *w.shape[1:]))
        if self.variant == "tokenfusion":
            x = x + x_res
        y = self.norm(x).mean(dim=1).view(B, T, C)
        if self.variant == "safuser":
            return y, torch.stack(attns).transpose(0, 1)
        return y


def erank_torch(x: torch.Tensor, rtol: float = 1e-4) -> torch.Tensor:
    """Effective-rank restatement (NOT reference code; parity unpinned) with
    torch.linalg.svdvals, differentiable -- used as the fp32 CPU baseline for
    the erank half of the metric and as an autograd cross-check."""
    s = torch.linalg.svdvals(x)
    keep = s > rtol * s.amax(dim=-1, keepdim=True)
    s = torch.where(keep, s, torch.zeros_like(s))
    p = s / s.sum(dim=-1, keepdim=True)
    plogp = torch.where(keep, p * torch.log(torch.where(keep, p, torch.ones_like(p))), torch.zeros_like(p))
    return torch.exp(-plogp.sum(dim=-1))



class PortCMFuserM(nn.Module):
    """M-modality generalisation of the tokenfusion fuser -- NOT reference code beyond M = 2 (PARITY UNPINNED).

    The reference reads ``M = len(modal_feats)`` (model/futr_safuser_tokenfusion.py:76) but hard-codes the keys
    'rgb' / 'depth' (:79) and a 2 x 2 mask (:77).  This restates the obvious extension r3d_b200 implements: modality m
    swaps its k = C // 4 lowest-score channels (eval-branch score, :49-50) for those of modality (m + 1) mod M; the
    streams are stacked to (B, T, M, C); the Block attends over the M tokens with the M x M -inf-diagonal mask
    (generate_cross_attention_mask(M), :68-72); outer residual, LayerNorm and the mean over the M tokens as :92-95.
    For M = 2 it equals PortCMFuser(variant='tokenfusion') -- tests/test_oracle_golden.py checks that."""

    def __init__(self, dim, depth=1, num_heads=4, mlp_ratio=4.0, qkv_bias=False):
        super().__init__()
        self.blocks = nn.ModuleList([PortBlock(dim, num_heads, mlp_ratio, qkv_bias) for _ in range(depth)])
        self.norm = nn.LayerNorm(dim)
        self.embd_drop = nn.Dropout(0.1)
        self.modality_token = nn.Parameter(torch.randn(1, 1, 1, dim))
        self.projection = nn.Linear(dim, dim)
        self.fusion_conv = nn.Conv2d(2, 1, kernel_size=1)

    def token_fusion(self, feats):
        M = len(feats)
        C = feats[0].shape[-1]
        k = C // 4
        idx = [torch.sort(f.detach().abs().mean(dim=(0, 1)), stable=True)[1][:k] for f in feats]
        outs = []
        for m in range(M):
            ex = feats[m].clone()
            ex[:, :, idx[m]] = feats[(m + 1) % M][:, :, idx[m]]
            outs.append(ex)
        self.last_indices = idx
        return torch.stack(outs, dim=2)

    def forward(self, modal_feats, mode="test"):
        feats = list(modal_feats.values())
        B, T, C = feats[0].shape
        M = len(feats)
        mask = torch.zeros(M, M, dtype=feats[0].dtype).masked_fill(torch.eye(M) == 1, float("-inf"))
        x = self.embd_drop(self.token_fusion(feats).reshape(B * T, M, C))
        x_res = x
        for blk in self.blocks:
            x, _ = blk(x, mask)
        x = x + x_res
        return self.norm(x).mean(dim=1).view(B, T, C)
