"""Generates the golden vectors under tests/golden/ by running the REFERENCE's own modules.

Run in the build container only (needs /root/reference; nothing on the GPU box reads it):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

What is recorded (inputs and weights are NOT stored -- they are regenerated bit-exactly
from seeds by ``svol_b200.synth``, which is numpy-only):

  head_<case>.npz      reference SVANet.forward outputs (fp32 and fp64 runs): logits / boxes of
                       every decoder layer, a strided sample of hs and of the last layer's mem.
  crit_<case>.npz      reference PerFrameMatcher / HungarianMatcher indices and SetCriterion
                       losses on synthetic predictions and on the head outputs above.
  lsap_cases.npz       scipy.optimize.linear_sum_assignment outputs (scipy is the reference's
                       third-party solver, matcher.py:8,93,158) on random / tied / inf / tall /
                       wide / empty cost matrices.
  post_<case>.npz      test.py:133-158 post-processing (scores, clamped xyxy, per-frame order).

Versions are written into every file.
"""
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
import scipy
import torch
from scipy.optimize import linear_sum_assignment

from lib.modeling.svanet import build_svanet            # noqa: E402  (reference)
from lib.modeling.loss import build_loss                # noqa: E402  (reference)
from lib.modeling.matcher import build_matcher          # noqa: E402  (reference)
from lib.utils.box_utils import box_cxcywh_to_xyxy      # noqa: E402  (reference)

from svol_b200 import synth                             # noqa: E402
from dataclasses import replace                         # noqa: E402

VERSIONS = np.array([f"torch={torch.__version__}", f"scipy={scipy.__version__}", f"numpy={np.__version__}"])
torch.set_num_threads(8)


def ref_head(cfg, sd_np, dtype):
    model = build_svanet(cfg.to_namespace())
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd_np.items()}, strict=True)
    return model.to(dtype).eval()


def run_head(cfg, batch, seed, padded, dtype):
    sd = synth.random_state_dict(cfg, seed)
    inp = synth.make_inputs(cfg, batch, seed, padded=padded)
    model = ref_head(cfg, sd, dtype)
    model.vis_mode = "on"                                   # returns (out, hs)   svanet.py:138-141
    t = lambda a: torch.from_numpy(a).to(dtype)
    with torch.no_grad():
        out, hs = model(t(inp["src_sketch"]), t(inp["src_sketch_mask"]), t(inp["src_video"]), t(inp["src_video_mask"]))
    logits = torch.stack([a["pred_logits"] for a in out["aux_outputs"]] + [out["pred_logits"]])
    boxes = torch.stack([a["pred_boxes"] for a in out["aux_outputs"]] + [out["pred_boxes"]])
    return out, logits.numpy(), boxes.numpy(), hs.numpy()


def head_case(name, cfg, batch, seed, padded):
    rec = {"versions": VERSIONS, "batch": batch, "seed": seed, "padded": padded}
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        out, logits, boxes, hs = run_head(cfg, batch, seed, padded, dt)
        rec[f"logits_{tag}"] = logits
        rec[f"boxes_{tag}"] = boxes
        rec[f"hs_sample_{tag}"] = hs[:, :, ::7, ::5].copy()
        if tag == "f32":
            out32 = out
    np.savez_compressed(os.path.join(HERE, f"head_{name}.npz"), **rec)
    print("head", name, rec["logits_f32"].shape, float(np.abs(rec["logits_f32"] - rec["logits_f64"]).max()))
    return out32


def indices_to_arrays(indices):
    pred = np.concatenate([i.numpy() for i, _ in indices])
    tgt = np.concatenate([j.numpy() for _, j in indices])
    counts = np.array([len(i) for i, _ in indices], np.int64)
    return pred, tgt, counts


def crit_case(name, cfg, batch, seed, outputs=None, max_per_frame=2, frame_mask=None):
    """Reference matcher indices + criterion losses on synthetic predictions (or on ``outputs``)."""
    targets_np = synth.make_targets(cfg, batch, seed, max_per_frame=max_per_frame, frame_mask=frame_mask)
    targets = synth.targets_to_torch(targets_np)
    if outputs is None:
        logits, boxes = synth.make_predictions(cfg, batch, seed)
        outputs = {"pred_logits": torch.from_numpy(logits[-1]), "pred_boxes": torch.from_numpy(boxes[-1]),
                   "aux_outputs": [{"pred_logits": torch.from_numpy(a), "pred_boxes": torch.from_numpy(b)}
                                   for a, b in zip(logits[:-1], boxes[:-1])]}
    ns = cfg.to_namespace()
    criterion = build_loss(ns).eval()
    matcher = build_matcher(ns)
    rec = {"versions": VERSIONS, "batch": batch, "seed": seed, "max_per_frame": max_per_frame}
    layers = [outputs] + list(outputs["aux_outputs"])
    for li, o in enumerate(layers):                          # li = 0 is the LAST decoder layer
        idx = matcher({"pred_logits": o["pred_logits"], "pred_boxes": o["pred_boxes"]}, targets)
        p, t, c = indices_to_arrays(idx)
        rec[f"pred_idx_{li}"], rec[f"tgt_idx_{li}"], rec[f"counts_{li}"] = p, t, c
    with torch.no_grad():
        losses = criterion(outputs, targets)
    for k, v in losses.items():
        rec["loss/" + k] = np.float32(float(v))
    rec["weight_keys"] = np.array(sorted(criterion.weight_dict.keys()))
    rec["weight_vals"] = np.array([criterion.weight_dict[k] for k in sorted(criterion.weight_dict.keys())], np.float64)
    np.savez_compressed(os.path.join(HERE, f"crit_{name}.npz"), **rec)
    print("crit", name, {k: round(float(v), 5) for k, v in losses.items()})


def post_case(name, cfg, batch, seed):
    logits, boxes = synth.make_predictions(cfg, batch, seed, layers=1)
    lg, bx = torch.from_numpy(logits[0]), torch.from_numpy(boxes[0])
    scores = torch.softmax(lg, -1)[..., 0]                   # test.py:133-134
    outs, orders = [], []
    for b in range(batch):
        xyxy = torch.clamp(box_cxcywh_to_xyxy(bx[b]), min=0, max=1)       # test.py:145
        preds = torch.cat([xyxy, scores[b][:, None]], dim=1)
        for chunk in preds.chunk(cfg.num_frames, dim=0):                   # test.py:153
            rows = list(chunk)
            order = sorted(range(len(rows)), key=lambda i: rows[i][4], reverse=True)   # test.py:157 (stable)
            outs.append(torch.stack([rows[i] for i in order]).numpy())
            orders.append(np.array(order, np.int64))
    qf = cfg.num_queries_per_frame
    np.savez_compressed(os.path.join(HERE, f"post_{name}.npz"), versions=VERSIONS, batch=batch, seed=seed,
                        sorted=np.stack(outs).reshape(batch, cfg.num_frames, qf, 5),
                        order=np.stack(orders).reshape(batch, cfg.num_frames, qf))
    print("post", name)


def lsap_cases():
    rng = np.random.RandomState(7)
    costs, shapes, rows, cols = [], [], [], []
    def add(c):
        r, k = linear_sum_assignment(c)
        costs.append(np.asarray(c, np.float64).ravel()); shapes.append(c.shape); rows.append(r); cols.append(k)
    for trial in range(400):
        nr, nc = rng.randint(1, 13), rng.randint(1, 13)
        kind = trial % 5
        if kind == 0:
            c = rng.rand(nr, nc).astype(np.float32)
        elif kind == 1:
            c = rng.randint(0, 3, size=(nr, nc)).astype(np.float64)        # heavy ties
        elif kind == 2:
            c = np.full((nr, nc), float(rng.randint(-2, 3)))               # constant
        elif kind == 3:
            c = rng.randint(0, 5, size=(nr, nc)).astype(np.float64)
            c[rng.rand(nr, nc) < 0.08] = np.inf
            try:
                linear_sum_assignment(c)
            except ValueError:
                continue
        else:
            c = (rng.standard_normal((nr, nc)) * 3).astype(np.float32)
        add(c)
    for shape in ((10, 1), (10, 2), (10, 3), (10, 10), (10, 12), (100, 50), (320, 64), (64, 320), (3, 0), (0, 4)):
        add(rng.rand(*shape).astype(np.float32))
    np.savez_compressed(os.path.join(HERE, "lsap_cases.npz"), versions=VERSIONS,
                        cost=np.concatenate(costs), shapes=np.array(shapes, np.int64),
                        rows=np.concatenate(rows).astype(np.int64), cols=np.concatenate(cols).astype(np.int64))
    print("lsap", len(shapes), "cases")


if __name__ == "__main__":
    C = synth.CONFIGS
    lsap_cases()
    # head forward
    out_tiny = head_case("tiny", C["tiny"], 3, 0, padded=False)
    out_tiny_pad = head_case("tiny_pad", C["tiny"], 3, 1, padded=True)
    out_c1a = head_case("C1a", C["C1a"], 2, 0, padded=False)
    out_c1b = head_case("C1b", C["C1b"], 2, 0, padded=False)
    out_c1b_pad = head_case("C1b_pad", C["C1b"], 2, 1, padded=True)
    # matcher + criterion on synthetic predictions
    crit_case("tiny", C["tiny"], 3, 0)
    crit_case("C2_b4", C["C2"], 4, 0)
    crit_case("C2_b32", C["C2"], 32, 1)
    crit_case("C2n4_b2", C["C2n4"], 2, 2)
    crit_case("C2_video", replace(C["C2"], matcher="video_matcher"), 4, 3)
    crit_case("C5_b2", C["C5"], 2, 0, max_per_frame=50)
    crit_case("C5_video", replace(C["C2"], matcher="video_matcher"), 2, 4, max_per_frame=10)
    # matcher + criterion on real head outputs (end-to-end chain)
    pad_mask = synth.make_inputs(C["C1b"], 2, 1, padded=True)["frame_mask"]
    crit_case("C1b_e2e", C["C1b"], 2, 0, outputs=out_c1b)
    crit_case("C1b_pad_e2e", C["C1b"], 2, 1, outputs=out_c1b_pad, frame_mask=pad_mask)
    post_case("C2_b4", C["C2"], 4, 0)
