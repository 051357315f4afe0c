"""Parameter containers of the cross-modal transformer (lib/modeling/cross_modal_transformer.py).

The modules reproduce the reference's parameter names, shapes and initialisation so reference
checkpoints load with ``strict=True`` (test.py:72-88); they carry no forward of their own -- the
computation is the launch plan in ``svol_b200/engine.py``.
"""
from __future__ import annotations

import copy

import torch
from torch import nn


class AttentionParams(nn.Module):
    """Same state_dict layout as nn.MultiheadAttention(d_model, nhead): packed ``in_proj_weight``
    (3d, d), ``in_proj_bias`` (3d), ``out_proj.{weight,bias}`` (cross_modal_transformer.py:88-97)."""

    def __init__(self, d_model: int, nhead: int):
        super().__init__()
        self.embed_dim, self.num_heads = d_model, nhead
        self.in_proj_weight = nn.Parameter(torch.empty(3 * d_model, d_model))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * d_model))
        self.out_proj = nn.Linear(d_model, d_model)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.constant_(self.out_proj.bias, 0.0)


class MLP(nn.Module):
    """fc1 -> GELU(erf) -> fc2 (cross_modal_transformer.py:163-179)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, activation="gelu"):
        super().__init__()
        if activation != "gelu":
            raise NotImplementedError("the fused epilogue implements the reference's live activation, GELU(erf)")
        self.fc1 = nn.Linear(in_features, hidden_features or in_features)
        self.fc2 = nn.Linear(hidden_features or in_features, out_features or in_features)


class CrossModalTransformerLayer(nn.Module):
    def __init__(self, d_model=512, nhead=8, dim_feedforward=2048, activation="gelu"):
        super().__init__()
        self.sketch_video_cross_attn = AttentionParams(d_model, nhead)
        self.norm1 = nn.LayerNorm(d_model)
        self.content_self_attn = AttentionParams(d_model, nhead)
        self.norm2 = nn.LayerNorm(d_model)
        self.mlp1 = MLP(d_model, dim_feedforward, activation=activation)
        self.norm3 = nn.LayerNorm(d_model)
        self.token_self_attn = AttentionParams(d_model, nhead)
        self.norm4 = nn.LayerNorm(d_model)
        self.content_token_cross_attn = AttentionParams(d_model, nhead)
        self.norm5 = nn.LayerNorm(d_model)
        self.mlp2 = MLP(d_model, dim_feedforward, activation=activation)
        self.norm6 = nn.LayerNorm(d_model)


class CrossModalTransformer(nn.Module):
    def __init__(self, d_model=512, nhead=8, num_layers=6, dim_feedforward=2048, activation="gelu"):
        super().__init__()
        layer = CrossModalTransformerLayer(d_model, nhead, dim_feedforward, activation)
        self.layers = nn.ModuleList([copy.deepcopy(layer) for _ in range(num_layers)])
        for p in self.parameters():                      # cross_modal_transformer.py:22-25
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        self.d_model, self.nhead, self.num_layers = d_model, nhead, num_layers


def build_cross_modal_transformer(args):
    """cross_modal_transformer.py:196-202 -- the FFN width is hard-wired to 2048 there."""
    return CrossModalTransformer(d_model=args.hidden_dim, nhead=args.nheads, num_layers=args.num_layers,
                                 dim_feedforward=2048)
