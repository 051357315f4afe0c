"""Round-2 golden vectors: the REFERENCE's own modules run at the BASELINE configurations' full sizes.

Run in the build container only (needs /root/reference; nothing on the GPU box reads it):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_r2.py [case ...]

  head_C2_b32.npz      reference SVANet.forward at the headline batch (BASELINE configs[1]: B=32, L=1568, Q=320,
                       2 layers, padded variant): logits / boxes of every decoder layer from an fp64 run (stored as
                       fp32; the fp32 run's distance to it is recorded as a scalar), PLUS the reference
                       PerFrameMatcher's indices on those outputs and, per (layer, video, frame), the cost gap between
                       the best and the second-best assignment of that frame's problem (matcher.py:85-96) -- what the
                       end-to-end index-agreement count needs (frames whose gap is below the float tolerance are
                       excluded AND counted).
  head_C2n4_b4.npz     same at 4 decoder layers (configs.py:121), B=4.
  head_C2_b4_spread.npz  C2 with the spread box head (below), B=4, with the matching record: predicted boxes that differ
                       from query to query give the assignment problems cost gaps well above the bf16 forward error,
                       so most frames must agree bit-exactly.
  head_C4_b1.npz       long clip (BASELINE configs[3]: T=128, L=6272, Q=1280), one pair whose last 5 frames are
                       masked (cross_modal_transformer.py:151-156 key_padding_mask).
  head_C1b_spread.npz  C1b with a box head scaled up so that the sigmoid boxes spread over (0.05, 0.95) -- at the
                       default initialisation every box lies in [0.47, 0.53] and a tolerance on post-sigmoid values
                       says little about the box MLP.

Inputs and weights are regenerated bit-exactly from seeds by ``svol_b200.synth`` (numpy only).
"""
import os
import sys
import time

os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
_argv = sys.argv[1:]
sys.argv = ["x"]

import numpy as np                                       # noqa: E402
import scipy                                             # noqa: E402
import torch                                             # noqa: E402
from scipy.optimize import linear_sum_assignment         # noqa: E402

from lib.modeling.svanet import build_svanet             # noqa: E402  (reference)
from lib.modeling.matcher import build_matcher           # noqa: E402  (reference)
from lib.utils.box_utils import box_cxcywh_to_xyxy, generalized_box_iou   # noqa: E402  (reference)

from svol_b200 import synth                              # noqa: E402

VERSIONS = np.array([f"torch={torch.__version__}", f"scipy={scipy.__version__}", f"numpy={np.__version__}"])
torch.set_num_threads(os.cpu_count() or 8)


def run_head(cfg, sd, inp, dtype):
    model = build_svanet(cfg.to_namespace())
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    model = model.to(dtype).eval()
    t = lambda a: torch.from_numpy(a).to(dtype)
    with torch.no_grad():
        out = model(t(inp["src_sketch"]), t(inp["src_sketch_mask"]), t(inp["src_video"]), t(inp["src_video_mask"]))
    logits = torch.stack([a["pred_logits"] for a in out["aux_outputs"]] + [out["pred_logits"]])
    boxes = torch.stack([a["pred_boxes"] for a in out["aux_outputs"]] + [out["pred_boxes"]])
    return logits, boxes


def second_best_gap(c):
    """Cost of the best assignment of the (rows x cols) problem and its distance to the best assignment that differs
    from it in at least one pair (forbid each optimal pair in turn and re-solve)."""
    r, k = linear_sum_assignment(c)
    best = float(c[r, k].sum())
    second = np.inf
    for i, j in zip(r, k):
        c2 = c.copy()
        c2[i, j] = np.inf
        try:
            r2, k2 = linear_sum_assignment(c2)
        except ValueError:
            continue
        second = min(second, float(c2[r2, k2].sum()))
    return best, second - best


def matching_record(cfg, logits32, boxes32, targets_np):
    """Reference PerFrameMatcher on every decoder layer's outputs + per-frame cost gaps from the reference's own cost
    formula (matcher.py:59-83), block-diagonal entries only, fp32 like the reference, gaps in fp64."""
    targets = synth.targets_to_torch(targets_np)
    matcher = build_matcher(cfg.to_namespace())
    NL, B = logits32.shape[:2]
    T, qf = cfg.num_frames, cfg.num_queries_per_frame
    rec = {}
    gaps = np.full((NL, B, T), np.inf, np.float64)
    for li in range(NL):
        idx = matcher({"pred_logits": logits32[li], "pred_boxes": boxes32[li]}, targets)
        rec[f"pred_idx_{li}"] = np.concatenate([i.numpy() for i, _ in idx])
        rec[f"tgt_idx_{li}"] = np.concatenate([j.numpy() for _, j in idx])
        rec[f"counts_{li}"] = np.array([len(i) for i, _ in idx], np.int64)
        prob = logits32[li].softmax(-1)
        for b in range(B):
            frames = list(targets[b]["bboxes"].values())
            for t, fr in enumerate(frames):
                if not fr:
                    continue
                tb = torch.stack([o["bbox"] for o in fr])
                ob = boxes32[li, b, t * qf:(t + 1) * qf]
                c = (cfg.set_cost_bbox * torch.cdist(ob, tb, p=1)
                     - cfg.set_cost_giou * generalized_box_iou(box_cxcywh_to_xyxy(ob), box_cxcywh_to_xyxy(tb))
                     - cfg.set_cost_class * prob[b, t * qf:(t + 1) * qf, :1])
                gaps[li, b, t] = second_best_gap(c.double().numpy())[1]
    rec["frame_gap"] = gaps
    return rec


def head_case(name, cfg, batch, seed, padded, mask_tail_frames=0, box_spread=0.0, with_matching=False):
    t0 = time.time()
    sd = synth.random_state_dict(cfg, seed, box_spread=box_spread)
    inp = synth.make_inputs(cfg, batch, seed, padded=padded)
    if mask_tail_frames:
        inp["src_video_mask"][-1, -mask_tail_frames * cfg.tokens_per_frame:] = 0
        inp["frame_mask"][-1, -mask_tail_frames:] = 0
    lg64, bx64 = run_head(cfg, sd, inp, torch.float64)
    lg32, bx32 = run_head(cfg, sd, inp, torch.float32)
    rec = {"versions": VERSIONS, "batch": batch, "seed": seed, "padded": padded, "mask_tail_frames": mask_tail_frames,
           "box_spread": box_spread,
           "logits_f64": lg64.float().numpy(), "boxes_f64": bx64.float().numpy(),
           "f32_vs_f64_logits": np.float64((lg32.double() - lg64).abs().max()),
           "f32_vs_f64_boxes": np.float64((bx32.double() - bx64).abs().max())}
    if with_matching:
        targets_np = synth.make_targets(cfg, batch, seed, frame_mask=inp["frame_mask"])
        rec.update(matching_record(cfg, lg32, bx32, targets_np))
        rec["logits_f32"], rec["boxes_f32"] = lg32.numpy(), bx32.numpy()
    np.savez_compressed(os.path.join(HERE, f"head_{name}.npz"), **rec)
    b = rec["boxes_f64"]
    print(f"head {name}: logits {rec['logits_f64'].shape} |max| {np.abs(rec['logits_f64']).max():.3f} boxes in "
          f"[{b.min():.3f}, {b.max():.3f}] f32-f64 {float(rec['f32_vs_f64_logits']):.2e} ({time.time() - t0:.0f} s)", flush=True)


CASES = {
    "C1b_spread": lambda: head_case("C1b_spread", synth.CONFIGS["C1b"], 2, 4, padded=True, box_spread=1.0),
    "C2n4_b4": lambda: head_case("C2n4_b4", synth.CONFIGS["C2n4"], 4, 9, padded=True, with_matching=True),
    "C2_b32": lambda: head_case("C2_b32", synth.CONFIGS["C2"], 32, 8, padded=True, with_matching=True),
    "C2_b4_spread": lambda: head_case("C2_b4_spread", synth.CONFIGS["C2"], 4, 12, padded=True, box_spread=1.0, with_matching=True),
    "C4_b1": lambda: head_case("C4_b1", synth.CONFIGS["C4"], 1, 11, padded=False, mask_tail_frames=5),
}

if __name__ == "__main__":
    for name in (_argv or list(CASES)):
        CASES[name]()
