"""GPU parity of the evaluation metrics: svol_eval_max_iou / svol_eval_average_precision behind
svol_b200.evaluate.SVOLEvaluator against the oracle (oracle/eval_oracle.py) and against the metric dictionaries the
reference's own eval_svol produced (tests/golden/eval_*.npz).  IoU maxima are compared bit for bit (float64, same
operation order), per-unit AP to 1e-12, the formatted metric dictionaries exactly."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import eval_oracle as ev
from svol_b200 import synth

pytestmark = pytest.mark.gpu


def _run(cfg, batch, seed, split=None):
    from svol_b200.evaluate import SVOLEvaluator, flatten_eval_targets
    from svol_b200.modeling import postprocess
    logits, boxes = synth.make_eval_predictions(cfg, batch, seed)
    inp = synth.make_inputs(cfg, batch, seed, padded=True)
    targets = synth.make_targets(cfg, batch, seed, frame_mask=inp["frame_mask"])
    evaluator = SVOLEvaluator(cfg.num_frames, cfg.num_queries_per_frame)
    posts = []
    for b0, b1 in (split or [(0, batch)]):
        post, _ = postprocess(torch.from_numpy(logits[b0:b1]).cuda(), torch.from_numpy(boxes[b0:b1]).cuda(), cfg.num_frames)
        evaluator.update(post, targets[b0:b1])
        posts.append(post.cpu().numpy().reshape(-1, cfg.num_queries_per_frame, 5))
    post = np.concatenate(posts)
    gt, gt_off, frame_off, frame_index = flatten_eval_targets(targets, cfg.num_frames)
    return evaluator, post[frame_index], gt, gt_off, frame_off


@pytest.mark.parametrize("name", ["C2_b4", "C2_b3"])
def test_eval_metrics_match_reference(name, golden_dir):
    gold = np.load(os.path.join(golden_dir, f"eval_{name}.npz"))
    cfg = synth.CONFIGS["C2"]
    evaluator, pred, gt, gt_off, frame_off = _run(cfg, int(gold["batch"]), int(gold["seed"]))
    assert np.array_equal(torch.cat(evaluator.max1).cpu().numpy(), ev.max_ious(pred, gt, gt_off, 1))
    assert np.array_equal(torch.cat(evaluator.max5).cpu().numpy(), ev.max_ious(pred, gt, gt_off, 5))
    ap_ref = np.stack([ev.average_precision_unit(pred, gt, gt_off, int(frame_off[v]), int(frame_off[v + 1]))
                       for v in range(len(frame_off) - 1)])
    assert np.abs(torch.cat(evaluator.ap).cpu().numpy() - ap_ref).max() < 1e-12
    assert evaluator.summary() == json.loads(str(gold["metrics"]))


def test_eval_accumulates_over_batches_and_long_units():
    """Two update() calls == one; a long-clip unit (T = 128: 1280 predictions per unit) against the oracle."""
    cfg = synth.CONFIGS["C2"]
    one, *_ = _run(cfg, 4, 0)
    two, *_ = _run(cfg, 4, 0, split=[(0, 1), (1, 4)])
    assert one.summary() == two.summary()
    cfg4 = synth.CONFIGS["C4"]
    evaluator, pred, gt, gt_off, frame_off = _run(cfg4, 2, 3)
    assert evaluator.summary() == ev.eval_svol(pred, gt, gt_off, frame_off)
