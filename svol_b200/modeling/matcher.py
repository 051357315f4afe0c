"""Hungarian matchers, drop-in for lib/modeling/matcher.py, solved on the GPU.

``PerFrameMatcher`` / ``HungarianMatcher`` keep the reference's constructor arguments, call signature
and return value (a list of ``(pred_idx, tgt_idx)`` CPU int64 tensors per video).  The cost blocks and
the assignment are computed by ``svol_match`` (one warp per frame / video problem); the only host work
is flattening ``targets`` once per batch.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Tuple

import torch
from torch import nn

from .. import _lib
from .targets import FlatTargets, TargetsCache


def fill_match_args(a, logits, boxes, src, NL, B, Q, K_pitch, ppv, rows, max_cols, per_frame, w_class, w_bbox, w_giou,
                    cost_ws, pred_idx, tgt_idx, status, mode=0, solver=0):
    """``src``: anything with tgt_boxes / tgt_off / match_off / cost_off / video_match_off / video_tgt_off tensors
    (a FlatTargets, or the static views of a criterion workspace)."""
    g = (lambda k: src[k]) if isinstance(src, dict) else (lambda k: getattr(src, k))
    a.logits, a.boxes, a.tgt_boxes = logits.data_ptr(), boxes.data_ptr(), g("tgt_boxes").data_ptr()
    a.tgt_off, a.match_off, a.cost_off = g("tgt_off").data_ptr(), g("match_off").data_ptr(), g("cost_off").data_ptr()
    a.cost_ws = cost_ws.data_ptr() if cost_ws is not None else None
    a.pred_idx, a.tgt_idx, a.status = pred_idx.data_ptr(), tgt_idx.data_ptr(), status.data_ptr()
    a.NL, a.B, a.Q = NL, B, Q
    a.problems_per_video, a.rows_per_problem, a.max_cols = ppv, rows, max_cols
    a.w_class, a.w_bbox, a.w_giou = float(w_class), float(w_bbox), float(w_giou)
    a.K = K_pitch
    a.video_match_off, a.video_tgt_off = g("video_match_off").data_ptr(), g("video_tgt_off").data_ptr()
    # global -> video-local target indices inside the call's finalize kernel.  PerFrameMatcher subtracts the minimum
    # matched global index (matcher.py:114-115); HungarianMatcher's columns are local to the video's split (:158).
    a.mode, a.solver, a.localize = mode, solver, (1 if per_frame else 2)
    a.order = g("order").data_ptr()            # problems with the most targets first
    return a


COST_SMEM_LIMIT = 96 * 1024          # matcher.cu: larger cost blocks are solved from the global workspace


def run_match(logits: torch.Tensor, boxes: torch.Tensor, flat: FlatTargets, w_class: float, w_bbox: float,
              w_giou: float, export_cost: bool = False, solver: int = 0, mode: int = 0, cost_ws: torch.Tensor = None):
    """logits [NL,B,Q,2], boxes [NL,B,Q,4] fp32 contiguous on the GPU.  Returns device tensors
    (pred_idx [NL,K] i64, tgt_idx [NL,K] i64 video-local, status [2] i32 (status[0] = this call's), cost_ws or None).
    Allocating convenience path (stand-alone matcher calls, tests); the criterion keeps a static workspace (loss.py)."""
    _lib.require_device()
    lib = _lib.get_lib()
    NL, B, Q = logits.shape[:3]
    dev = logits.device
    need_ws = export_cost or mode != 0 or flat.rows_per_problem * flat.max_cols * 4 > COST_SMEM_LIMIT
    if cost_ws is None and need_ws:
        cost_ws = torch.empty((NL, max(flat.cost_total, 1)), device=dev, dtype=torch.float32)
    pred_idx = torch.empty((NL, flat.K), device=dev, dtype=torch.int64)
    tgt_idx = torch.empty((NL, flat.K), device=dev, dtype=torch.int64)
    status = torch.zeros(2, device=dev, dtype=torch.int32)
    a = fill_match_args(_lib.MatchArgs(), logits, boxes, flat, NL, B, Q, flat.K, flat.problems_per_video,
                        flat.rows_per_problem, flat.max_cols, flat.per_frame, w_class, w_bbox, w_giou, cost_ws, pred_idx,
                        tgt_idx, status, mode, solver)
    _lib.check(lib.svol_match(C.byref(a), _lib.stream_ptr()), "match")
    return pred_idx, tgt_idx, status, cost_ws


def _to_index_list(pred_idx: torch.Tensor, tgt_idx: torch.Tensor, flat: FlatTargets) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    p, t = pred_idx.cpu(), tgt_idx.cpu()          # the one device sync of the matcher API (matcher.py:86)
    off = flat.h_video_match_off
    return [(p[off[b]:off[b + 1]].clone(), t[off[b]:off[b + 1]].clone()) for b in range(flat.B)]


def _check_status(status: torch.Tensor) -> None:
    s = int(status[0].item())
    if s & 1:
        raise ValueError("matrix contains invalid numeric entries")      # scipy's message for NaN / -inf costs
    if s & 2:
        raise ValueError("cost matrix is infeasible")


class _MatcherBase(nn.Module):
    def __init__(self, cost_class: float = 1, cost_bbox: float = 1, cost_giou: float = 1):
        super().__init__()
        self.cost_class, self.cost_bbox, self.cost_giou = cost_class, cost_bbox, cost_giou
        self.foreground_label = 0
        assert cost_class != 0 or cost_bbox != 0 or cost_giou != 0, "all costs cant be 0"
        self._cache = TargetsCache()

    def _flat(self, targets, device, num_queries) -> FlatTargets:
        raise NotImplementedError

    @torch.no_grad()
    def forward(self, outputs, targets):
        logits = outputs["pred_logits"].detach().float().contiguous()[None]
        boxes = outputs["pred_boxes"].detach().float().contiguous()[None]
        flat = self._flat(targets, logits.device, boxes.shape[2])
        pred_idx, tgt_idx, status, _ = run_match(logits, boxes, flat, self.cost_class, self.cost_bbox, self.cost_giou)
        _check_status(status)
        return _to_index_list(pred_idx[0], tgt_idx[0], flat)


class PerFrameMatcher(_MatcherBase):
    """matcher.py:12-119: one ``num_queries_per_frame x n_f`` assignment per frame."""

    def __init__(self, cost_class: float = 1, cost_bbox: float = 1, cost_giou: float = 1, num_frames: int = 32,
                 num_queries_per_frame: int = 10):
        super().__init__(cost_class, cost_bbox, cost_giou)
        self.num_frames = num_frames
        self.num_queries_per_frame = num_queries_per_frame

    def _flat(self, targets, device, num_queries):
        assert num_queries == self.num_frames * self.num_queries_per_frame          # matcher.py:56
        return self._cache.get(targets, device, True, self.num_frames, num_queries, self.num_queries_per_frame)


class HungarianMatcher(_MatcherBase):
    """matcher.py:122-159: one ``Q x n_v`` assignment per video."""

    def _flat(self, targets, device, num_queries):
        return self._cache.get(targets, device, False, 0, num_queries, 0)


def build_matcher(args):
    """matcher.py:162-177."""
    if args.matcher == "per_frame_matcher":
        return PerFrameMatcher(cost_bbox=args.set_cost_bbox, cost_giou=args.set_cost_giou, cost_class=args.set_cost_class,
                               num_frames=args.num_frames, num_queries_per_frame=args.num_queries_per_frame)
    elif args.matcher == "video_matcher":
        return HungarianMatcher(cost_bbox=args.set_cost_bbox, cost_giou=args.set_cost_giou, cost_class=args.set_cost_class)
    raise NotImplementedError
