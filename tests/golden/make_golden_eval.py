"""Generates tests/golden/eval_C2_b4.npz by running the REFERENCE's own evaluation (lib/evaluate/eval.py: eval_svol)
on synthetic results built the way test.py:133-170 builds them (softmax score, clamped xyxy, per-frame sort, 4-decimal
rounding, ground truth converted with box_cxcywh_to_xyxy).  Build container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_eval.py
"""
import json
import os
import sys

os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.argv = ["x"]

import numpy as np
import torch
import torch.nn.functional as F

from lib.evaluate.eval import eval_svol                     # noqa: E402  (reference)
from lib.utils.box_utils import box_cxcywh_to_xyxy          # noqa: E402  (reference)

from svol_b200 import synth                                 # noqa: E402


def build_results(cfg, batch, seed):
    """test.py:133-170 on synthetic predictions that overlap the targets (so that mAP / recall are non-trivial)."""
    logits, boxes = synth.make_eval_predictions(cfg, batch, seed)
    targets = synth.targets_to_torch(synth.make_targets(cfg, batch, seed, frame_mask=synth.make_inputs(cfg, batch, seed, padded=True)["frame_mask"]))
    lg, bx = torch.from_numpy(logits), torch.from_numpy(boxes)
    scores = F.softmax(lg, -1)[..., 0]
    results = []
    for target, b, s in zip(targets, bx, scores):
        frame_idxs = list(target["bboxes"].keys())
        b = torch.clamp(box_cxcywh_to_xyxy(b), min=0, max=1)
        preds = torch.cat([b, s[:, None]], dim=1).chunk(cfg.num_frames, dim=0)
        for preds_per_frame, fidx in zip(preds, frame_idxs):
            sorted_preds = sorted(preds_per_frame, key=lambda x: x[4], reverse=True)
            sorted_preds = [[float(f"{e:.4f}") for e in row] for row in sorted_preds]
            gt_boxes = [{"track_id": o["track_id"], "bbox": box_cxcywh_to_xyxy(o["bbox"]).tolist()} for o in target["bboxes"][fidx]]
            results.append(dict(video=target["video"], sketch=target["sketch"], shape=target["size"], frame=fidx,
                                gt_boxes=gt_boxes, pred_boxes=sorted_preds))
    return results


if __name__ == "__main__":
    import logging
    cfg = synth.CONFIGS["C2"]
    for name, batch, seed in (("C2_b4", 4, 0), ("C2_b3", 3, 7)):
        results = build_results(cfg, batch, seed)
        metrics = eval_svol(results, verbose=False, logger=logging.getLogger("x"))
        np.savez_compressed(os.path.join(HERE, f"eval_{name}.npz"), batch=batch, seed=seed, metrics=json.dumps(metrics),
                            versions=np.array([f"torch={torch.__version__}", f"numpy={np.__version__}"]))
        print(name, json.dumps(metrics)[:300])
