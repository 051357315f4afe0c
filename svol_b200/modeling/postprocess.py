"""Inference post-processing of test.py:133-158 as one kernel."""
from __future__ import annotations

import torch

from .. import _lib


@torch.no_grad()
def postprocess(pred_logits: torch.Tensor, pred_boxes: torch.Tensor, num_frames: int):
    """pred_logits (B,Q,2), pred_boxes (B,Q,4) cxcywh -> (sorted (B,T,q_f,5) fp32 rows
    ``[x0,y0,x1,y1,score]`` ordered by foreground score within each frame, order (B,T,q_f) int32)."""
    _lib.require_device()
    B, Q = pred_boxes.shape[:2]
    qf = Q // num_frames
    lg = pred_logits.detach().float().contiguous()
    bx = pred_boxes.detach().float().contiguous()
    out = torch.empty((B, Q, 5), device=lg.device, dtype=torch.float32)
    order = torch.empty((B, Q), device=lg.device, dtype=torch.int32)
    _lib.check(_lib.get_lib().svol_postprocess(lg.data_ptr(), bx.data_ptr(), out.data_ptr(), order.data_ptr(), B, Q, qf,
                                               _lib.stream_ptr()), "postprocess")
    return out.view(B, num_frames, qf, 5), order.view(B, num_frames, qf)
