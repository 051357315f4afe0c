"""Positional encoding modules of the drop-in head (lib/modeling/position_encoding.py).

Only the sine variant is live in the reference ('trainable' passes constructor arguments that do
not exist, position_encoding.py:104-108; 'learned' is a 2-D image embedding).  The module is
parameter-free; on the CUDA path the table is produced by ``svol_posenc_sine`` inside the launch
plan, this class only exists so ``SVANet`` has the reference's attribute layout and so the table can
be requested on its own.
"""
from __future__ import annotations

import math

import torch
from torch import nn

from .. import _lib


class PositionEmbeddingSine(nn.Module):
    def __init__(self, num_pos_feats=64, temperature=10000, normalize=False, scale=None):
        super().__init__()
        if scale is not None and normalize is False:
            raise ValueError("normalize should be True if scale is passed")
        if temperature != 10000 or not normalize or scale not in (None, 2 * math.pi):
            raise NotImplementedError("the CUDA kernel implements temperature=10000, normalize=True, scale=2*pi")
        self.num_pos_feats = num_pos_feats
        self.temperature = temperature
        self.normalize = normalize
        self.scale = 2 * math.pi

    @torch.no_grad()
    def forward(self, x, mask):
        """x: (B, L, d) (only its device is used); mask: (B, L), nonzero = valid.  Returns (B, L, num_pos_feats) fp32."""
        assert mask is not None
        _lib.require_device()
        B, L = mask.shape
        m = mask.to(device=x.device, dtype=torch.float32).contiguous()
        pos = torch.empty((B, L, self.num_pos_feats), device=x.device, dtype=torch.float32)
        _lib.check(_lib.get_lib().svol_posenc_sine(m.data_ptr(), pos.data_ptr(), B, L, self.num_pos_feats,
                                                   _lib.stream_ptr()), "posenc_sine")
        return pos


def build_position_encoding(args):
    """position_encoding.py:101-129 (sine only)."""
    out = []
    for kind in (args.sketch_position_embedding, args.video_position_embedding):
        if kind != "sine":
            raise ValueError(f"not supported {kind} (the reference's 'trainable' / 'learned' variants are dead code)")
        out.append(PositionEmbeddingSine(args.hidden_dim, normalize=True))
    return tuple(out)
