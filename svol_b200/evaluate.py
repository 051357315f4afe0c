"""Evaluation metrics of the reference (lib/evaluate/eval.py: SVOL-mAP, SVOL-R1 / R5, mIoU@R1 / R5) on the GPU
(SURVEY 8f-3).  The reference builds a Python list of per-frame result dicts (test.py:133-170), writes / re-reads it
as JSONL and evaluates it with numpy loops over Python objects (8 worker processes for the mAP); here the per-frame
score-sorted predictions that ``svol_postprocess`` already produced stay on the device and two kernels compute, per
batch, the best IoU of every ground-truth box among the frame's top-1 / top-5 predictions and the VOC-style average
precision of every (video, sketch) unit at the 10 IoU thresholds.  ``SVOLEvaluator`` accumulates those small arrays over
the evaluation set and formats the reference's metric dictionary.

All arithmetic that decides a comparison (IoU vs threshold, greedy matching order) is done in float64 in the
reference's operation order on the 4-decimal-rounded predictions (test.py:161), so the counts are bit-exact.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np
import torch

from . import _lib

IOU_THDS_AP = [float(f"{e:.2f}") for e in np.linspace(0.5, 0.95, 10)]
IOU_THDS_RECALL = [float(f"{e:.2f}") for e in np.linspace(0.1, 0.9, 9)]


def flatten_eval_targets(targets: Sequence[dict], num_frames: int):
    """Ground truth of the evaluated frames as flat arrays (host).  A video contributes its first len(bboxes) frames
    (test.py:141,153: ``zip(preds, frame_idxs)``); boxes become xyxy in float32 like ``box_cxcywh_to_xyxy(bbox)``
    (test.py:163).  Returns gt (S,4) f32, gt_off (F+1) i32, frame_off (V+1) i32, frame_index (F) i32 into B*T frames."""
    gt, gt_off, frame_off, frame_index = [], [0], [0], []
    for b, t in enumerate(targets):
        for i, frame in enumerate(t["bboxes"].values()):
            for o in frame:
                c = np.asarray(o["bbox"], np.float32)
                hw, hh = np.float32(0.5) * c[2], np.float32(0.5) * c[3]
                gt.append([c[0] - hw, c[1] - hh, c[0] + hw, c[1] + hh])
            gt_off.append(len(gt))
            frame_index.append(b * num_frames + i)
        frame_off.append(len(frame_index))
    return (np.asarray(gt, np.float32).reshape(-1, 4), np.asarray(gt_off, np.int32), np.asarray(frame_off, np.int32),
            np.asarray(frame_index, np.int32))


class SVOLEvaluator:
    """Accumulates the evaluation of successive batches; ``summary()`` returns eval_svol's dictionary (eval.py:102-117)."""

    def __init__(self, num_frames: int, q_per_frame: int):
        self.num_frames, self.q_per_frame = num_frames, q_per_frame
        self.max1: List[torch.Tensor] = []
        self.max5: List[torch.Tensor] = []
        self.ap: List[torch.Tensor] = []

    def update(self, post: torch.Tensor, targets: Sequence[dict]) -> None:
        """post [B, Q, 5]: the per-frame score-sorted (x0, y0, x1, y1, score) rows of ``svol_postprocess``."""
        _lib.require_device()
        dev = post.device
        gt, gt_off, frame_off, frame_index = flatten_eval_targets(targets, self.num_frames)
        S, F, V = gt.shape[0], frame_index.shape[0], frame_off.shape[0] - 1
        d = lambda a: torch.from_numpy(a).to(dev)
        d_gt, d_goff, d_foff, d_fidx = d(gt), d(gt_off), d(frame_off), d(frame_index)
        max1 = torch.empty(S, device=dev, dtype=torch.float64)
        max5 = torch.empty(S, device=dev, dtype=torch.float64)
        ap = torch.empty((V, len(IOU_THDS_AP)), device=dev, dtype=torch.float64)
        lib, P = _lib.get_lib(), _lib.ptr
        post = post.contiguous().float()
        _lib.check(lib.svol_eval_max_iou(P(post), P(d_fidx), P(d_gt), P(d_goff), F, S, self.q_per_frame, P(max1), P(max5),
                                         _lib.stream_ptr()), "eval_max_iou")
        max_frames = int(np.diff(frame_off).max())
        max_gt = int((gt_off[frame_off[1:]] - gt_off[frame_off[:-1]]).max())
        _lib.check(lib.svol_eval_average_precision(P(post), P(d_fidx), P(d_gt), P(d_goff), P(d_foff), V, self.q_per_frame, max_frames,
                                                   max_gt, P(ap), _lib.stream_ptr()), "eval_average_precision")
        self.max1.append(max1); self.max5.append(max5); self.ap.append(ap)

    def summary(self) -> Dict[str, object]:
        ap = torch.cat(self.ap).cpu().numpy()
        m1, m5 = torch.cat(self.max1).cpu().numpy(), torch.cat(self.max5).cpu().numpy()
        ap_thds = ap.mean(0)
        m_ap = {str(t): float(f"{100 * v:.2f}") for t, v in zip(IOU_THDS_AP, ap_thds)}
        m_ap["average"] = float(f"{100 * np.mean(ap_thds):.2f}")
        rec = lambda m: {str(t): float(f"{np.mean(m >= t) * 100:.2f}") for t in IOU_THDS_RECALL}
        return {"SVOL-mAP": m_ap, "SVOL-R1": rec(m1), "SVOL-R5": rec(m5),
                "mIoU@R1": float(f"{np.mean(m1) * 100:.2f}"), "mIoU@R5": float(f"{np.mean(m5) * 100:.2f}")}
