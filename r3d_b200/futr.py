"""``FUTR`` around the B200 fuser path -- the caller on the far side of the hot path.

Mirrors ``model/futr_safuser_tokenfusion.py:99-239``: same constructor, same ``forward(inputs, depth_features, mode)``
contract and output dict, and the same ``state_dict`` names (``input_embed``, ``depth_projection``, ``depth_layernorm``,
``fuser.*``, ``transformer.encoder.* / decoder.*``, ``query_embed``, ``fc``, ``fc_len``, ``fc_seg``, ``fc_l3``,
``l3_attention``, ``query_attention``, ``pos_embedding``), so reference checkpoints load with ``strict=True``.

What runs where:
  * RGB embedding + ReLU, depth projection + LayerNorm + ReLU (tokenfusion.py:179-197): tcgen05 GEMMs with the channel
    score as a by-product (r3d_b200.embed);
  * the fuser (tokenfusion.py:199): r3d_b200.CMFuser, fed the score by-products (no separate score pass);
  * the decoder (model/extras/transformer.py:126,152-196,281-330; the encoder is bypassed there, :77-78) and the heads
    (tokenfusion.py:220-232): 8 queries cross-attending over T keys -- tiny next to the fuser, left on
    torch's ``nn.MultiheadAttention`` / ``nn.Linear`` (library code on the same device, not a CPU path).
The unused encoder layers are instantiated only so that checkpoints load.
"""
from __future__ import annotations

import copy
from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn

from .embed import DepthEmbed, RGBEmbed
from .fuser import CMFuser


class _EncoderLayer(nn.Module):
    """Parameter container for the bypassed encoder (transformer.py:216-238 names)."""

    def __init__(self, d_model, nhead, dim_feedforward, dropout):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)


class _DecoderLayer(nn.Module):
    """Post-norm DETR decoder layer (transformer.py:277-330, forward_post)."""

    def __init__(self, d_model, nhead, dim_feedforward, dropout):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.multihead_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.norm3 = nn.LayerNorm(d_model)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.dropout3 = nn.Dropout(dropout)

    def forward(self, tgt, memory, memory_key_padding_mask, pos, query_pos):
        q = tgt + query_pos                      # the reference feeds the position-added tensor as value too (:289-291)
        tgt = self.norm1(tgt + self.dropout1(self.self_attn(q, q, value=q)[0]))
        mem = memory + pos
        tgt2 = self.multihead_attn(query=tgt + query_pos, key=mem, value=mem, key_padding_mask=memory_key_padding_mask)[0]
        tgt = self.norm2(tgt + self.dropout2(tgt2))
        tgt2 = self.linear2(self.dropout(F.relu(self.linear1(tgt))))
        return self.norm3(tgt + self.dropout3(tgt2))


class _Stack(nn.Module):
    def __init__(self, layer, n, norm=None):
        super().__init__()
        self.layers = nn.ModuleList([copy.deepcopy(layer) for _ in range(n)])
        self.norm = norm


class Transformer(nn.Module):
    """transformer.py:20-128 with the FUTR call pattern: memory = src (encoder bypassed), decoder over the queries."""

    def __init__(self, d_model, nhead, num_encoder_layers, num_decoder_layers, dim_feedforward, dropout=0.1):
        super().__init__()
        self.encoder = _Stack(_EncoderLayer(d_model, nhead, dim_feedforward, dropout), num_encoder_layers)
        self.decoder = _Stack(_DecoderLayer(d_model, nhead, dim_feedforward, dropout), num_decoder_layers,
                              nn.LayerNorm(d_model))
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

    def forward(self, src, tgt, mask, query_embed, pos_embed):
        out = tgt
        for layer in self.decoder.layers:
            out = layer(out, src, mask, pos_embed, query_embed)
        return src, self.decoder.norm(out)


class _SinusoidTable(nn.Module):
    """``pos_enc`` / ``pos_enc_depth`` of the reference (model/extras/position.py:15-27): a sinusoidal buffer
    ``pos_table`` (1, 3000, C) that FUTR.forward never adds (tokenfusion.py:185 is commented out); kept because it is
    part of the checkpoint."""

    def __init__(self, d_model: int, max_len: int = 3000):
        super().__init__()
        import math
        position = torch.arange(max_len).unsqueeze(1)
        div = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
        pe = torch.zeros(1, max_len, d_model)
        pe[0, :, 0::2] = torch.sin(position * div)
        pe[0, :, 1::2] = torch.cos(position * div)
        self.register_buffer("pos_table", pe)


class FUTR(nn.Module):
    """Drop-in for ``model/futr_safuser_tokenfusion.py::FUTR`` (i3d_transcript inputs)."""

    def __init__(self, n_class, hidden_dim, src_pad_idx, device, args, n_query=8, n_head=8, num_encoder_layers=6,
                 num_decoder_layers=6, query_num=49, depth_hw: int = 224 * 224, fuser_kw: Optional[dict] = None):
        super().__init__()
        self.src_pad_idx = src_pad_idx
        self.query_pad_idx = query_num - 1
        self.device = device
        self.hidden_dim = hidden_dim
        self.n_query = n_query
        self.args = args
        self._rgb = RGBEmbed(args.input_dim, hidden_dim)
        self.input_embed = self._rgb.input_embed                       # tokenfusion.py:111 (same Parameter objects)
        self.transformer = Transformer(hidden_dim, n_head, num_encoder_layers, num_decoder_layers, hidden_dim * 4)
        self.l3_attention = nn.MultiheadAttention(hidden_dim, n_head, batch_first=True)
        self.query_attention = nn.MultiheadAttention(hidden_dim, n_head, batch_first=True)
        self.query_embed = nn.Embedding(self.n_query, hidden_dim)
        self.fuser = CMFuser(dim=hidden_dim, depth=1, num_heads=n_head, **(fuser_kw or {}))
        if getattr(args, "seg", False):
            self.fc_seg = nn.Linear(hidden_dim, n_class)
            nn.init.xavier_uniform_(self.fc_seg.weight)
        if getattr(args, "anticipate", False):
            self.fc = nn.Linear(hidden_dim, n_class)
            nn.init.xavier_uniform_(self.fc.weight)
            self.fc_len = nn.Linear(hidden_dim, 1)
            nn.init.xavier_uniform_(self.fc_len.weight)
        self.fc_l3 = nn.Linear(hidden_dim, query_num)
        self.pos_embedding = nn.Parameter(torch.zeros(1, args.max_pos_len, hidden_dim))
        nn.init.xavier_uniform_(self.pos_embedding)
        self.pos_enc = _SinusoidTable(hidden_dim)
        self.pos_enc_depth = _SinusoidTable(hidden_dim)
        self._depth = DepthEmbed(depth_hw, hidden_dim)
        self.depth_projection = self._depth.depth_projection             # tokenfusion.py:143
        self.depth_layernorm = self._depth.depth_layernorm               # :147

    def state_dict(self, *a, **k):
        sd = super().state_dict(*a, **k)
        return type(sd)((n, v) for n, v in sd.items() if not (n.startswith("_rgb.") or n.startswith("_depth.")))

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        sd = dict(state_dict)
        for own, ref in (("_rgb.input_embed.", "input_embed."), ("_depth.depth_projection.", "depth_projection."),
                         ("_depth.depth_layernorm.", "depth_layernorm.")):
            for suffix in ("weight", "bias"):
                if ref + suffix in sd:
                    sd[own + suffix] = sd[ref + suffix]
        return super().load_state_dict(sd, strict=strict, **kw)

    def forward(self, inputs, depth_features, mode="train", epoch=0, idx=0):
        # the reference variants disagree on tuple vs tensor inputs outside training: accept both
        if isinstance(inputs, (tuple, list)):
            src, src_label = inputs
        else:
            src, src_label = inputs, None
        mask = None
        if mode == "train" and src_label is not None:
            mask = (src_label == self.src_pad_idx).to(src.device)         # get_pad_mask, tokenfusion.py:168,242
        B, S, _ = src.shape
        src = self._rgb(src)                                              # relu(input_embed(src)), :179-183
        depth = self._depth(depth_features.reshape(B, S, -1))             # relu(LN(depth_projection(.))), :194-197
        fused = self.fuser({"rgb": src, "depth": depth}, mode,
                           score_parts=(self._rgb.last_score, self._depth.last_score))          # :199
        fused = fused.transpose(0, 1)                                     # b t c -> t b c
        pos = self.pos_embedding[:, :S].repeat(B, 1, 1).transpose(0, 1)
        query = self.query_embed.weight.unsqueeze(0).repeat(B, 1, 1).transpose(0, 1)
        tgt = torch.zeros_like(query)
        mem, hs = self.transformer(fused, tgt, mask, query, pos)
        hs, mem = hs.transpose(0, 1), mem.transpose(0, 1)
        out = {}
        if getattr(self.args, "anticipate", False):
            out["action"] = self.fc(hs)
            out["duration"] = self.fc_len(hs).squeeze(2)
        if getattr(self.args, "seg", False):
            out["seg"] = self.fc_seg(mem)
        return out
