"""Drop-in ``CMFuser`` for R3D's FUTR models, backed by the sm_100a kernels.

Keeps the reference constructor, ``forward(modal_feats, mode)`` signature, public
``token_fusion`` / ``generate_cross_attention_mask`` and every ``state_dict`` name
(reference: model/futr_safuser_tokenfusion.py:17-97, ..._vary.py:17-87,
..._batchnormalization.py:17-107, futr_safuser_depth.py:17-64), so
``FUTR.__init__`` can assign it to ``self.fuser`` unchanged
(model/futr_safuser_tokenfusion.py:120,199) and reference checkpoints load.

Differences from the reference, all documented in DESIGN.md:
  * ties in the channel score go to the lower channel index (north_star);
    the reference's ``torch.topk`` tie order is an implementation artefact;
  * ``mode == 'train'`` of the tokenfusion variant has a data-independent
    constant score: the wasted ``autograd.grad`` is not replayed,
    the tie rule is applied directly (channels 0..k-1);
  * the 2x2 masked self-attention is evaluated through its exact closed form
    (each modality token attends only to the other one), so W_q / W_k
    are never multiplied -- their gradients are exactly zero in the reference too;
  * no mask tensor is built and nothing is copied to 'cuda' per call.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from ._lib import R3DError

VARIANTS = ("tokenfusion", "vary", "batchnorm", "safuser")


class Attention(nn.Module):
    """Parameter container matching model/extras/transformerblock.py:7-17
    (``qkv`` 3C x C, ``proj``); evaluated in closed form by :class:`Block`."""

    def __init__(self, dim, num_heads=8, qkv_bias=False):
        super().__init__()
        self.num_heads = num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)


class MLP(nn.Module):
    """model/extras/transformerblock.py:79-93 (``mlp.0``, ``mlp.2`` names kept)."""

    def __init__(self, in_features, hidden_features):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(in_features, hidden_features), nn.GELU(),
                                 nn.Linear(hidden_features, in_features), nn.Dropout(0.0))

    def forward(self, x):
        return self.mlp(x)


def _ln(mod: nn.LayerNorm, x: torch.Tensor) -> torch.Tensor:
    """nn.LayerNorm through the hand-written row kernels (ops.layer_norm) whenever they cover the shape; the module
    object stays an nn.LayerNorm so that state_dict names and checkpoints are the reference's."""
    C = x.shape[-1]
    if mod.elementwise_affine and mod.weight is not None and mod.bias is not None \
            and tuple(mod.normalized_shape) == (C,) and ops.layer_norm_supported(x, C):
        return ops.layer_norm(x, mod.weight, mod.bias, mod.eps)
    return mod(x)


class Block(nn.Module):
    """model/extras/transformerblock.py:118-135 specialised to the fuser's use:
    two tokens per row and a -inf diagonal mask.  softmax([-inf, a]) == [0, 1]
    exactly, so attention output for token m is proj(V[1 - m])."""

    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False):
        super().__init__()
        self.dim = dim
        self.norm1 = nn.LayerNorm(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = MLP(dim, int(dim * mlp_ratio))

    def forward(self, x: torch.Tensor, attn_mask=None):
        C = self.dim
        fc1, fc2 = self.mlp.mlp[0], self.mlp.mlp[2]
        if x.shape[1] != 2:
            # M >= 3 modality tokens: a real masked softmax over the other M - 1 tokens (csrc/multi_kernels.cu)
            hd = C // self.attn.num_heads
            if not (x.is_cuda and x.shape[1] <= 4 and C % self.attn.num_heads == 0 and hd % 32 == 0 and hd <= 256
                    and C % 8 == 0 and ops.layer_norm_supported(x, C)):
                raise R3DError("the M-modality Block needs CUDA tensors, 3 <= M <= 4, head_dim a multiple of 32 (<= 256)")
            return ops.fused_block_multi(x, self.norm1, self.attn.qkv, self.attn.proj, self.norm2, fc1, fc2,
                                         self.attn.num_heads), None
        if ops.fused_block_supported(x, C, fc1.out_features) and self.norm1.elementwise_affine \
                and self.norm2.elementwise_affine and fc1.bias is not None and fc2.bias is not None \
                and self.attn.proj.bias is not None:
            # the whole Block on hand-written kernels: tcgen05 GEMMs with fused epilogues, LayerNorm kernels
            return ops.fused_block(x, self.norm1, self.attn.qkv, self.attn.proj, self.norm2, fc1, fc2), None
        h = _ln(self.norm1, x)
        w_v = self.attn.qkv.weight[2 * C:]
        b_v = None if self.attn.qkv.bias is None else self.attn.qkv.bias[2 * C:]
        v = F.linear(h, w_v, b_v)                       # (R, 2, C): only the V third of qkv
        p = self.attn.proj(v)                           # per-token projection commutes with the token swap
        if p.dtype == x.dtype and ops.swap_add_supported(x):
            x = ops.swap_add(x, p)                      # token m <- projected V of token 1-m, fused with the residual
        else:
            x = x + p.flip(1)
        x = x + self.mlp(_ln(self.norm2, x))
        return x, None


class CMFuser(nn.Module):
    """Rank-enhancing token fuser (SA-Fuser with channel exchange).

    Reference-compatible positional arguments; ``variant`` picks which of the four
    reference files is reproduced.

    ``score_scope='local'`` (default) is the reference's behaviour under
    ``nn.DataParallel`` (main_utkinects.py:129): every replica / rank scores its own
    shard and ``forward`` issues NO collective, so ranks may call it independently
    (rank-0-only validation, uneven last batches).  ``score_scope='global'`` is opt-in:
    it all-reduces the (2, C) score sums over ``process_group`` so that the selection
    equals the single-process result on the concatenated batch -- EVERY rank of the
    group must then call ``forward`` the same number of times, or the collective hangs.

    ``select_axis='channel'`` (default) is what the reference ships (tokenfusion.py:33-66 exchanges CHANNELS chosen by a
    batch-global score).  ``select_axis='token'`` is the form the paper prose describes (README.md:13; no
    reference code, parity unpinned): per sample, the T // 4 tokens with the lowest spectral informativeness
    (``ops.token_scores``, a by-product of the effective-rank chain) are replaced by the other modality's tokens;
    ``last_erank`` then holds the per-sample effective ranks (rgb, depth) of the call.
    """

    def __init__(self, dim, depth=1, num_heads=4, mlp_ratio=4.0, qkv_bias=False, *, variant: str = "tokenfusion",
                 score_scope: str = "local", process_group=None, select_axis: str = "channel"):
        super().__init__()
        if variant not in VARIANTS:
            raise ValueError(f"variant must be one of {VARIANTS}")
        if score_scope not in ("global", "local"):
            raise ValueError("score_scope must be 'global' or 'local'")
        if select_axis not in ("channel", "token"):
            raise ValueError("select_axis must be 'channel' or 'token'")
        if select_axis == "token" and variant != "tokenfusion":
            raise ValueError("select_axis='token' is defined for the swap variant ('tokenfusion') only")
        self.dim = dim
        self.num_heads = num_heads
        self.variant = variant
        self.score_scope = score_scope
        self.process_group = process_group
        self.select_axis = select_axis
        self.last_erank = None
        # optional per-sample statistic (e.g. the effective ranks of the step) whose SUM rides in the packed score
        # buffer [sum|rgb| | sum|depth| | sum statistic | rows], so that global scope needs ONE all-reduce for both;
        # last_packed is that buffer after the (possible) all-reduce
        self.statistic_extra = None
        self.last_packed = None
        self.blocks = nn.ModuleList([Block(dim, num_heads, mlp_ratio, qkv_bias) for _ in range(depth)])
        self.norm = nn.LayerNorm(dim)
        self.embd_drop = nn.Dropout(0.1)
        self.modality_token = nn.Parameter(torch.randn(1, 1, 1, dim))
        self.projection = nn.Linear(dim, dim)
        if variant != "safuser":
            self.fusion_conv = nn.Conv2d(in_channels=2, out_channels=1, kernel_size=1)
        if variant == "vary":
            self.alpha = nn.Parameter(torch.ones(1, 1, dim))
        elif variant == "batchnorm":
            self.alpha = nn.Parameter(torch.rand(1, 1, dim))
            self.bn_rgb = nn.BatchNorm1d(dim, affine=True)
            self.bn_depth = nn.BatchNorm1d(dim, affine=True)
        self.last_indices = None

    # ----- reference API -------------------------------------------------------------
    @staticmethod
    def generate_cross_attention_mask(sz):
        """Kept for API compatibility (tokenfusion.py:68-72); the kernels never use it."""
        mask = torch.eye(sz)
        return mask.masked_fill(mask == 1, float("-inf"))

    def k_for(self, C: int) -> int:
        return max(0, int(C * 0.1)) if self.variant == "batchnorm" else C // 4

    def _maybe_allreduce(self, packed: torch.Tensor) -> torch.Tensor:
        """Global scope: all-reduce the packed [sum|rgb| (C) || sum|depth| (C) || sum erank || rows] buffer (written
        on the device by one kernel -- no pageable host copy, no sync)."""
        import torch.distributed as dist
        if self.score_scope == "global" and dist.is_available() and dist.is_initialized() \
                and dist.get_world_size(self.process_group) > 1:
            dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=self.process_group)
        return packed

    def select_channels(self, rgb: torch.Tensor, depth: torch.Tensor, mode: str, score_parts=None):
        """score -> bottom-k.  Returns (idx_r, idx_d) int64 (k,) each.  score_parts: optional pair of
        embed.ScoreParts (the |x| column-sum partials the producers of rgb / depth emitted); the score pass is then
        skipped."""
        B, T, C = rgb.shape
        k = self.k_for(C)
        if self.variant == "batchnorm":
            score = torch.stack([self.bn_rgb.weight.detach().abs().float(),
                                 self.bn_depth.weight.detach().abs().float()])
        elif self.variant == "tokenfusion" and mode == "train":
            # tokenfusion.py:40-45: constant score 1/(B*T*C) for every channel -> all ties
            # -> tie rule = lowest index first
            idx = torch.arange(k, device=rgb.device, dtype=torch.int64)
            return idx, idx.clone()
        else:
            if score_parts is not None and self.variant in ("tokenfusion", "vary"):
                from .embed import pack_score_parts
                packed = pack_score_parts(score_parts[0], score_parts[1], self.statistic_extra)
            else:
                packed = ops.channel_score_packed(rgb.detach(), depth.detach(), self.statistic_extra)
            packed = self._maybe_allreduce(packed)
            self.last_packed = packed
            idx = ops.bottomk_packed(packed, k)        # the mean = sums / rows is formed inside the kernel
            return idx[0], idx[1]
        idx = ops.bottomk(score, k)
        return idx[0], idx[1]

    def token_fusion(self, rgb_feats: torch.Tensor, depth_feats: torch.Tensor, mode: str, score_parts=None) -> torch.Tensor:
        """(B,T,C) x2 -> (B,T,2,C).  tokenfusion.py:33-66 / vary.py:34-59 / batchnorm.py:38-77."""
        if self.variant == "safuser":
            raise R3DError("the safuser variant has no token_fusion (futr_safuser_depth.py has none)")
        if not rgb_feats.is_cuda:
            raise R3DError("r3d_b200.CMFuser runs on CUDA tensors only (no CPU fallback)")
        if self.select_axis == "token":
            T = rgb_feats.shape[1]
            s_r, er_r = ops.token_scores(rgb_feats, return_erank=True)
            s_d, er_d = ops.token_scores(depth_feats, return_erank=True)
            idx_r, idx_d = ops.bottomk(s_r, T // 4), ops.bottomk(s_d, T // 4)          # (B, k) each, ties -> lower index
            self.last_indices, self.last_erank = (idx_r, idx_d), (er_r, er_d)
            return ops.token_exchange(rgb_feats, depth_feats, idx_r, idx_d)
        idx_r, idx_d = self.select_channels(rgb_feats, depth_feats, mode, score_parts)
        self.last_indices = (idx_r, idx_d)
        if self.variant == "tokenfusion":
            return ops.exchange(rgb_feats, depth_feats, idx_r, idx_d, None, ops.BLEND_SWAP)
        if self.variant == "vary":
            return ops.exchange(rgb_feats, depth_feats, idx_r, idx_d, self.alpha, ops.BLEND_SCALE)
        return self._token_fusion_bn(rgb_feats, depth_feats, idx_r, idx_d)

    def _token_fusion_bn(self, rgb, depth, idx_r, idx_d):
        bn_r, bn_d = self.bn_rgb, self.bn_depth
        use_batch = bn_r.training or bn_r.running_mean is None
        if use_batch:
            stats = ops.bn_batch_stats(rgb.detach(), depth.detach())          # (2, 3, C)
            mean, var = stats[:, 0], stats[:, 1]
            if bn_r.training and bn_r.track_running_stats:
                with torch.no_grad():
                    for i, bn in enumerate((bn_r, bn_d)):
                        bn.num_batches_tracked += 1
                        mom = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
                        bn.running_mean.mul_(1 - mom).add_(stats[i, 0].to(bn.running_mean.dtype), alpha=mom)
                        bn.running_var.mul_(1 - mom).add_(stats[i, 2].to(bn.running_var.dtype), alpha=mom)
        else:
            mean = torch.stack([bn_r.running_mean, bn_d.running_mean]).float()
            var = torch.stack([bn_r.running_var, bn_d.running_var]).float()
        return ops.token_fusion_bn(rgb, depth, self.alpha, bn_r.weight, bn_r.bias, bn_d.weight, bn_d.bias, mean, var,
                                   idx_r, idx_d, bn_r.eps, use_batch)

    def _forward_multi(self, modal_feats: Dict[str, torch.Tensor]):
        """M >= 3 modalities (BASELINE.json configs[4]: rgb + depth + gaze): modality m swaps its k lowest-score
        channels for those of modality (m + 1) mod M, the Block attends over the M tokens (M x M -inf-diagonal mask),
        then LayerNorm and the mean over the M tokens.  The reference hard-codes M = 2 (tokenfusion.py:77,79); this is
        its extension, parity unpinned (oracle/torch_port.py:PortCMFuserM)."""
        if self.variant != "tokenfusion" or self.select_axis != "channel":
            raise R3DError("more than two modalities are defined for the channel-exchange 'tokenfusion' variant only")
        feats = list(modal_feats.values())
        B, T, C = feats[0].shape
        M = len(feats)
        k = self.k_for(C)
        sums = []
        for i in range(0, M, 2):                                   # the score kernel takes two tensors per launch
            a, b = feats[i].detach(), feats[min(i + 1, M - 1)].detach()
            s2 = ops.channel_score_sums(a, b)
            sums.append(s2 if i + 1 < M else s2[:1])
        packed = torch.cat([torch.cat(sums).reshape(-1), feats[0].new_zeros(1, dtype=torch.float32),
                            torch.full((1,), float(B * T), dtype=torch.float32, device=feats[0].device)])
        import torch.distributed as dist
        if self.score_scope == "global" and dist.is_available() and dist.is_initialized() \
                and dist.get_world_size(self.process_group) > 1:
            dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=self.process_group)
        score = packed[:M * C].reshape(M, C) / packed[M * C + 1]
        idx = ops.bottomk(score, k)                                 # (M, k), ties -> lower index
        self.last_indices = tuple(idx[m] for m in range(M))
        x = ops.exchange_multi(feats, idx)                          # (B, T, M, C)
        x = self.embd_drop(x.view(B * T, M, C))
        x_res = x
        for blk in self.blocks:
            x, _ = blk(x)
        x = x + x_res                                               # tokenfusion.py:92
        y = ops.token_mean(_ln(self.norm, x).contiguous())          # norm, then the mean over the M tokens (:93-95)
        return y.view(B, T, C)

    def forward(self, modal_feats: Dict[str, torch.Tensor], mode: Optional[str] = None, score_parts=None):
        if len(modal_feats) > 2:
            return self._forward_multi(modal_feats)
        rgb, depth = modal_feats["rgb"], modal_feats["depth"]
        B, T, C = rgb.shape
        M = len(modal_feats)
        if self.variant == "safuser":
            # futr_safuser_depth.py:43-49: concat + learned modality token (no exchange)
            x = torch.stack([rgb, depth], dim=2) + self.modality_token
        else:
            x = self.token_fusion(rgb, depth, "test" if mode is None else mode, score_parts)
        x = self.embd_drop(x.view(B * T, 2, C))
        x_res = x
        for blk in self.blocks:
            x, _ = blk(x)
        if self.variant == "tokenfusion":
            x = x + x_res                                     # tokenfusion.py:92
        nm = self.norm
        if nm.elementwise_affine and x.dim() == 3 and x.shape[1] == 2 and ops.layer_norm_supported(x, C):
            y = ops.layer_norm_mean2(x, nm.weight, nm.bias, nm.eps).view(B, T, C)   # norm + mean over the 2 tokens
        else:
            y = _ln(nm, x).mean(dim=1).view(B, T, C)
        if self.variant == "safuser":
            # attention weights are the constant [[0,1],[1,0]]:
            # (B, depth, T, heads, 2, 2) as futr_safuser_depth.py:64 returns
            w = torch.tensor([[0.0, 1.0], [1.0, 0.0]], device=y.device, dtype=y.dtype)
            attn = w.expand(B, len(self.blocks), T, self.num_heads, 2, 2)
            return y, attn
        return y


# Same-named entry points as the four reference files, for `from ... import CMFuser` swaps.
class TokenFusionCMFuser(CMFuser):
    def __init__(self, dim, depth=1, num_heads=4, mlp_ratio=4.0, qkv_bias=False, **kw):
        super().__init__(dim, depth, num_heads, mlp_ratio, qkv_bias, variant="tokenfusion", **kw)


class VaryCMFuser(CMFuser):
    def __init__(self, dim, depth=1, num_heads=4, mlp_ratio=4.0, qkv_bias=False, **kw):
        super().__init__(dim, depth, num_heads, mlp_ratio, qkv_bias, variant="vary", **kw)


class BatchNormCMFuser(CMFuser):
    def __init__(self, dim, depth=1, num_heads=4, mlp_ratio=4.0, qkv_bias=False, **kw):
        super().__init__(dim, depth, num_heads, mlp_ratio, qkv_bias, variant="batchnorm", **kw)


class SAFuser(CMFuser):
    def __init__(self, dim, depth=1, num_heads=4, mlp_ratio=4.0, qkv_bias=False, **kw):
        super().__init__(dim, depth, num_heads, mlp_ratio, qkv_bias, variant="safuser", **kw)
