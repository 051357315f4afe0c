"""Host-side marshalling of the reference's nested ``targets`` structure into flat device arrays.

The reference re-walks the nested dicts once per decoder layer in the matcher (matcher.py:62-70) and
again in the box loss (loss.py:78-85).  Here the walk happens once per batch; the result is cached on
the batch object's identity so the criterion, the matcher and repeated layers share it.
Flatten order = video -> frame (dict insertion order) -> instance, exactly as the reference.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence

import numpy as np
import torch


@dataclass
class FlatTargets:
    B: int
    S: int                       # total target boxes
    K: int                       # total matched pairs per layer
    P: int                       # assignment problems per layer
    problems_per_video: int
    rows_per_problem: int
    max_cols: int
    cost_total: int              # floats of cost workspace per layer
    tgt_boxes: torch.Tensor      # [S,4] f32 (device)
    tgt_off: torch.Tensor        # [P+1] i32
    match_off: torch.Tensor      # [P+1] i32
    cost_off: torch.Tensor       # [P+1] i64
    video_tgt_off: torch.Tensor  # [B+1] i32
    video_match_off: torch.Tensor  # [B+1] i32
    match_video: torch.Tensor    # [K] i32
    h_video_match_off: np.ndarray
    per_frame: bool


def _walk(targets: Sequence[dict]):
    boxes, per_video, per_frame_counts = [], [], []
    for t in targets:
        per_frame_counts.extend(int(n) for n in t["num_boxes_per_frame"])
        cnt = 0
        for frame in t["bboxes"].values():
            for inst in frame:
                bx = inst["bbox"]
                boxes.append(bx.detach().cpu().numpy() if isinstance(bx, torch.Tensor) else np.asarray(bx))
                cnt += 1
        per_video.append(cnt)
    if not boxes:
        raise ValueError("targets contain no boxes (the dataset guarantees >= 1 per video, svol_dataset.py:272)")
    return np.stack(boxes).astype(np.float32).reshape(-1, 4), per_video, per_frame_counts


def flatten_targets(targets: Sequence[dict], device, per_frame: bool, num_frames: int, num_queries: int,
                    q_per_frame: int) -> FlatTargets:
    boxes, per_video, per_frame_counts = _walk(targets)
    B = len(targets)
    if per_frame:
        if len(per_frame_counts) != B * num_frames:
            raise ValueError("'num_boxes_per_frame' must hold exactly num_frames entries per video (svol_dataset.py:267-271)")
        if sum(per_frame_counts) != boxes.shape[0]:
            raise ValueError("'num_boxes_per_frame' does not add up to the number of boxes under 'bboxes'")
        cols = np.asarray(per_frame_counts, np.int64)
        rows, ppv = q_per_frame, num_frames
    else:
        cols = np.asarray(per_video, np.int64)
        rows, ppv = num_queries, 1
    P = cols.shape[0]
    tgt_off = np.zeros(P + 1, np.int64); tgt_off[1:] = np.cumsum(cols)
    matched = np.minimum(cols, rows)
    match_off = np.zeros(P + 1, np.int64); match_off[1:] = np.cumsum(matched)
    cost_off = np.zeros(P + 1, np.int64); cost_off[1:] = np.cumsum(cols * rows)
    video_tgt_off = np.zeros(B + 1, np.int64); video_tgt_off[1:] = np.cumsum(per_video)
    video_match_off = match_off[::ppv].copy()
    K = int(match_off[-1])
    match_video = np.repeat(np.arange(B, dtype=np.int32), np.diff(video_match_off))

    ints = np.concatenate([tgt_off, match_off, video_tgt_off, video_match_off, match_video]).astype(np.int32)
    h_int = torch.from_numpy(ints)
    h_box = torch.from_numpy(boxes)
    h_cost = torch.from_numpy(cost_off)
    if device.type == "cuda":
        h_int, h_box, h_cost = h_int.pin_memory(), h_box.pin_memory(), h_cost.pin_memory()
    d_int = h_int.to(device, non_blocking=True)
    n1, n2, n3, n4 = P + 1, 2 * (P + 1), 2 * (P + 1) + B + 1, 2 * (P + 1) + 2 * (B + 1)
    return FlatTargets(
        B=B, S=int(boxes.shape[0]), K=K, P=P, problems_per_video=ppv, rows_per_problem=rows,
        max_cols=int(cols.max()), cost_total=int(cost_off[-1]),
        tgt_boxes=h_box.to(device, non_blocking=True), tgt_off=d_int[:n1], match_off=d_int[n1:n2],
        cost_off=h_cost.to(device, non_blocking=True), video_tgt_off=d_int[n2:n3], video_match_off=d_int[n3:n4],
        match_video=d_int[n4:], h_video_match_off=video_match_off, per_frame=per_frame)


class TargetsCache:
    """Remembers the flattening of the most recent ``targets`` list (keyed by object identity, length
    and device) -- the criterion calls into it once per forward, the standalone matcher once per call."""

    def __init__(self):
        self._key = None
        self._val = None
        self._ref = None

    def get(self, targets, device, per_frame, num_frames, num_queries, q_per_frame) -> FlatTargets:
        key = (id(targets), len(targets), str(device), per_frame, num_frames, num_queries, q_per_frame)
        if key != self._key or self._ref is not targets:
            self._val = flatten_targets(targets, device, per_frame, num_frames, num_queries, q_per_frame)
            self._key, self._ref = key, targets
        return self._val
