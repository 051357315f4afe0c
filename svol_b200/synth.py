"""Deterministic synthetic workloads for the SVOL hot path (numpy only).

Everything the parity tests, the golden-vector generator and ``bench.py`` feed into
the head / matcher / criterion comes from here, so the CUDA path, the oracle and the
imported reference all see bit-identical inputs and weights for a given seed.

Shapes follow the reference:
  * head inputs       -- ``SVANet.forward`` (lib/modeling/svanet.py:65-82)
  * ``targets`` list  -- ``SVOLDataset.__getitem__`` (lib/dataset/svol_dataset.py:234-288)
  * state_dict keys   -- ``SVANet.__init__`` (lib/modeling/svanet.py:17-63) and
                         ``CrossModalTransformerLayer.__init__``
                         (lib/modeling/cross_modal_transformer.py:86-100)
The distributions are the ones SURVEY.md section 8(d) fixes.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict
from typing import Dict, List, Tuple

import numpy as np


@dataclass(frozen=True)
class HeadConfig:
    """The fields of the reference's ``args`` namespace that the hot path reads
    (svanet.py:184-200, cross_modal_transformer.py:196-202, matcher.py:162-177,
    loss.py:192-213)."""
    hidden_dim: int = 256
    nheads: int = 8
    num_layers: int = 2
    num_queries: int = 320
    num_queries_per_frame: int = 10
    num_frames: int = 32
    tokens_per_frame: int = 49          # ResNet-34 7x7 grid; 1 for the ViT-CLS setting
    input_vid_dim: int = 512
    input_skch_dim: int = 512
    n_input_proj: int = 2
    dim_feedforward: int = 2048         # hard-wired in build_cross_modal_transformer
    input_dropout: float = 0.4
    aux_loss: bool = True
    matcher: str = "per_frame_matcher"
    set_cost_bbox: float = 5.0
    set_cost_giou: float = 1.0
    set_cost_class: float = 2.0
    eos_coef: float = 0.1

    @property
    def video_len(self) -> int:
        return self.num_frames * self.tokens_per_frame

    def to_namespace(self):
        import argparse
        ns = argparse.Namespace(**asdict(self))
        ns.use_sketch_pos = True
        ns.vis_mode = None
        ns.sketch_position_embedding = "sine"
        ns.video_position_embedding = "sine"
        ns.bbox_type = "cxcywh"
        ns.sketch_head = "svanet"
        return ns


# Named workloads (BASELINE.json configs / SURVEY.md section 8d).
CONFIGS: Dict[str, HeadConfig] = {
    # literal lib/configs.py defaults: ViT-CLS tokens, 4 layers
    "C1a": HeadConfig(num_layers=4, tokens_per_frame=1, input_vid_dim=768, input_skch_dim=768),
    # QuickDraw script setting (train_quickdraw.sh:18-28)
    "C1b": HeadConfig(num_layers=2),
    "C2": HeadConfig(num_layers=2),
    "C2n4": HeadConfig(num_layers=4),
    # 4x frames long-clip stress
    "C4": HeadConfig(num_layers=2, num_frames=128, num_queries=1280),
    # many-query / many-target matcher stress
    "C5": HeadConfig(num_layers=2, num_queries=3200, num_queries_per_frame=100),
    # tiny shapes for fast CPU tests
    "tiny": HeadConfig(num_layers=2, num_frames=4, tokens_per_frame=5, num_queries=12,
                       num_queries_per_frame=3, input_vid_dim=64, input_skch_dim=64,
                       dim_feedforward=2048),
}


def _uniform(rng, shape, bound):
    return rng.uniform(-bound, bound, size=shape).astype(np.float32)


def _xavier(rng, shape):
    fan_out, fan_in = shape
    return _uniform(rng, shape, math.sqrt(6.0 / (fan_in + fan_out)))


def _linear_default(rng, out_f, in_f):
    b = 1.0 / math.sqrt(in_f)
    return _uniform(rng, (out_f, in_f), b), _uniform(rng, (out_f,), b)


def random_state_dict(cfg: HeadConfig, seed: int = 0, perturb: bool = True, box_spread: float = 0.0
                      ) -> Dict[str, np.ndarray]:
    """Random weights under the reference's state_dict key names and shapes.

    Distributions mirror the reference's initialisers (xavier_uniform on the transformer
    matrices, cross_modal_transformer.py:22-25; torch defaults elsewhere).  With
    ``perturb`` the LayerNorm affine parameters and the attention biases, which the
    reference initialises to exactly 1/0, are jittered so that a kernel that drops one
    of them fails parity instead of passing by accident.  ``box_spread`` > 0 scales the box head (every layer of
    ``bbox_embed``, after all random draws, so the other tensors of a seed are unchanged) until the sigmoid boxes
    cover most of (0, 1): at the default initialisation every box lies within 0.03 of 0.5.
    """
    rng = np.random.RandomState(1000 + seed)
    d, ff = cfg.hidden_dim, cfg.dim_feedforward
    sd: Dict[str, np.ndarray] = {}

    def ln(prefix, n):
        if perturb:
            sd[prefix + ".weight"] = (1.0 + 0.1 * rng.standard_normal(n)).astype(np.float32)
            sd[prefix + ".bias"] = (0.1 * rng.standard_normal(n)).astype(np.float32)
        else:
            sd[prefix + ".weight"] = np.ones(n, np.float32)
            sd[prefix + ".bias"] = np.zeros(n, np.float32)

    for name, din in (("input_video_proj", cfg.input_vid_dim), ("input_sketch_proj", cfg.input_skch_dim)):
        dims = [din] + [d] * cfg.n_input_proj
        for i in range(cfg.n_input_proj):
            ln(f"{name}.{i}.LayerNorm", dims[i])
            w, b = _linear_default(rng, dims[i + 1], dims[i])
            sd[f"{name}.{i}.net.1.weight"], sd[f"{name}.{i}.net.1.bias"] = w, b

    sd["query_embed.weight"] = rng.standard_normal((cfg.num_queries, d)).astype(np.float32)
    for i, (o, k) in enumerate(((d, d), (d, d), (4, d))):
        sd[f"bbox_embed.layers.{i}.weight"], sd[f"bbox_embed.layers.{i}.bias"] = _linear_default(rng, o, k)
    sd["class_embed.weight"], sd["class_embed.bias"] = _linear_default(rng, 2, d)
    sd["class_head.weight"], sd["class_head.bias"] = _linear_default(rng, 2, d)   # unused by forward

    for li in range(cfg.num_layers):
        p = f"transformer.layers.{li}"
        for attn in ("sketch_video_cross_attn", "content_self_attn", "token_self_attn",
                     "content_token_cross_attn"):
            sd[f"{p}.{attn}.in_proj_weight"] = _xavier(rng, (3 * d, d))
            sd[f"{p}.{attn}.out_proj.weight"] = _xavier(rng, (d, d))
            if perturb:
                sd[f"{p}.{attn}.in_proj_bias"] = (0.05 * rng.standard_normal(3 * d)).astype(np.float32)
                sd[f"{p}.{attn}.out_proj.bias"] = (0.05 * rng.standard_normal(d)).astype(np.float32)
            else:
                sd[f"{p}.{attn}.in_proj_bias"] = np.zeros(3 * d, np.float32)
                sd[f"{p}.{attn}.out_proj.bias"] = np.zeros(d, np.float32)
        for n in range(1, 7):
            ln(f"{p}.norm{n}", d)
        for m in ("mlp1", "mlp2"):
            sd[f"{p}.{m}.fc1.weight"] = _xavier(rng, (ff, d))
            sd[f"{p}.{m}.fc1.bias"] = _uniform(rng, (ff,), 1.0 / math.sqrt(d))
            sd[f"{p}.{m}.fc2.weight"] = _xavier(rng, (d, ff))
            sd[f"{p}.{m}.fc2.bias"] = _uniform(rng, (d,), 1.0 / math.sqrt(ff))
    if box_spread > 0:
        for i in range(3):
            sd[f"bbox_embed.layers.{i}.weight"] = sd[f"bbox_embed.layers.{i}.weight"] * np.float32(1.0 + 1.6 * box_spread)
    return sd


def make_inputs(cfg: HeadConfig, batch: int, seed: int = 0, padded: bool = False,
                relu_features: bool = True) -> Dict[str, np.ndarray]:
    """Head inputs: frame-token features, one sketch feature per pair, float {0,1} masks.

    ``padded`` masks the last U{1..8} frames of ~25% of the videos (always video 0), which
    is the only way the reference's key_padding_mask (cross_modal_transformer.py:154) and
    the cumsum-based positional encoding (position_encoding.py:58-62) get exercised.
    """
    rng = np.random.RandomState(2000 + seed)
    L = cfg.video_len
    vid = rng.standard_normal((batch, L, cfg.input_vid_dim)).astype(np.float32)
    if relu_features:
        vid = np.maximum(vid, 0.0)
    skch = rng.standard_normal((batch, 1, cfg.input_skch_dim)).astype(np.float32)
    frame_mask = np.ones((batch, cfg.num_frames), np.float32)
    if padded:
        for b in range(batch):
            if b == 0 or rng.rand() < 0.25:
                cut = int(rng.randint(1, min(8, cfg.num_frames - 1) + 1))
                frame_mask[b, cfg.num_frames - cut:] = 0.0
    return {
        "src_sketch": skch,
        "src_sketch_mask": np.ones((batch, 1), np.float32),
        "src_video": vid,
        # model.py:21-22 -- per-frame mask repeated over the frame's tokens
        "src_video_mask": np.repeat(frame_mask, cfg.tokens_per_frame, axis=1),
        "frame_mask": frame_mask,
    }


def make_targets(cfg: HeadConfig, batch: int, seed: int = 0, max_per_frame: int = 2,
                 frame_mask: np.ndarray | None = None) -> List[dict]:
    """``targets`` in the reference's nested schema (svol_dataset.py:234-288), with numpy
    boxes; :func:`targets_to_torch` converts the leaves for the reference / the drop-in API.

    Boxes are cxcywh, non-degenerate and inside the image.  Frames that are masked out
    carry no boxes and (like the dataset's zero padding, svol_dataset.py:267-271) no dict
    entry, so ``num_boxes_per_frame`` is zero-padded to T while ``bboxes`` has fewer keys.
    """
    rng = np.random.RandomState(3000 + seed)
    T = cfg.num_frames
    out = []
    for b in range(batch):
        n_valid = T if frame_mask is None else int(frame_mask[b].sum())
        counts = rng.randint(0, max_per_frame + 1, size=n_valid)
        if counts.sum() == 0:
            counts[rng.randint(0, n_valid)] = 1     # svol_dataset.py:272 guarantees >= 1 box
        bboxes = {}
        for t in range(n_valid):
            frame = []
            for k in range(int(counts[t])):
                cxcy = rng.uniform(0.2, 0.8, size=2)
                wh = rng.uniform(0.05, 0.35, size=2)
                frame.append({"track_id": k,
                              "bbox": np.concatenate([cxcy, wh]).astype(np.float32)})
            bboxes[str(3 * t)] = frame
        nbpf = [0] * T
        for i, fr in enumerate(bboxes.values()):
            nbpf[i] = len(fr)
        out.append({
            "video": f"synthetic_{seed}_{b}", "size": [224, 224], "sketch": f"sketch_{seed}_{b}",
            "category": "synthetic", "track_ids": list(range(int(counts.max()))),
            "total_boxes": int(sum(nbpf)), "num_boxes_per_frame": nbpf, "bboxes": bboxes,
        })
    return out


def make_predictions(cfg: HeadConfig, batch: int, seed: int = 0, layers: int | None = None
                     ) -> Tuple[np.ndarray, np.ndarray]:
    """Synthetic head outputs for matcher / criterion tests that bypass the transformer:
    logits ~ N(0,1) of shape (layers,B,Q,2) and cxcywh boxes in (0,1) of shape (layers,B,Q,4)."""
    rng = np.random.RandomState(4000 + seed)
    n = cfg.num_layers if layers is None else layers
    logits = rng.standard_normal((n, batch, cfg.num_queries, 2)).astype(np.float32)
    cxcy = rng.uniform(0.15, 0.85, size=(n, batch, cfg.num_queries, 2))
    wh = rng.uniform(0.03, 0.4, size=(n, batch, cfg.num_queries, 2))
    boxes = np.concatenate([cxcy, wh], axis=-1).astype(np.float32)
    return logits, boxes


def make_eval_predictions(cfg: HeadConfig, batch: int, seed: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """Last-layer predictions for the evaluation-metric tests: logits (B,Q,2) ~ N(0,1) and cxcywh boxes (B,Q,4) of which
    roughly a third are jittered copies of the frame's ground-truth boxes (same seed as :func:`make_targets` with the
    padded frame mask), so that recall / mAP take non-trivial values; the rest are random boxes."""
    rng = np.random.RandomState(6000 + seed)
    inp = make_inputs(cfg, batch, seed, padded=True)
    targets = make_targets(cfg, batch, seed, frame_mask=inp["frame_mask"])
    logits = rng.standard_normal((batch, cfg.num_queries, 2)).astype(np.float32)
    cxcy = rng.uniform(0.15, 0.85, size=(batch, cfg.num_queries, 2))
    wh = rng.uniform(0.03, 0.4, size=(batch, cfg.num_queries, 2))
    boxes = np.concatenate([cxcy, wh], axis=-1).astype(np.float32)
    qf = cfg.num_queries_per_frame
    for b, t in enumerate(targets):
        for fi, frame in enumerate(t["bboxes"].values()):
            for o in frame:
                for _ in range(3):
                    q = fi * qf + int(rng.randint(0, qf))
                    boxes[b, q] = (np.asarray(o["bbox"]) * (1.0 + 0.07 * rng.standard_normal(4))).clip(0.02, 0.98).astype(np.float32)
                    logits[b, q, 0] += 2.0
    return logits, boxes


def make_upstream_grads(cfg: HeadConfig, batch: int, seed: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """Synthetic gradients w.r.t. the stacked head outputs (logits (layers,B,Q,2), boxes (layers,B,Q,4)) for the
    backward parity tests that bypass the matcher: N(0, 1e-2), the scale of d(mean loss)/d(output) times ~1e2."""
    rng = np.random.RandomState(5000 + seed)
    gl = (1e-2 * rng.standard_normal((cfg.num_layers, batch, cfg.num_queries, 2))).astype(np.float32)
    gb = (1e-2 * rng.standard_normal((cfg.num_layers, batch, cfg.num_queries, 4))).astype(np.float32)
    return gl, gb


def targets_to_torch(targets: List[dict]) -> List[dict]:
    """Deep-copies ``targets`` with every ``'bbox'`` leaf as a CPU float32 torch tensor,
    which is what the reference's matcher stacks (matcher.py:64-70)."""
    import torch
    out = []
    for t in targets:
        t2 = dict(t)
        t2["bboxes"] = {k: [{"track_id": o["track_id"], "bbox": torch.from_numpy(np.array(o["bbox"]))}
                            for o in fr] for k, fr in t["bboxes"].items()}
        out.append(t2)
    return out


def algorithmic_flops_per_pair(cfg: HeadConfig) -> float:
    """SURVEY.md section 8(d) formula: multiply-add = 2 FLOPs, dead V projection of the
    sketch->video attention excluded, softmax/LN/GELU not counted."""
    L, Q, d, ff, D = cfg.video_len, cfg.num_queries, cfg.hidden_dim, cfg.dim_feedforward, cfg.input_vid_dim
    per_layer = (2 * L * d * d + 2 * d * d + 2 * L * d
                 + 8 * L * d * d + 4 * L * L * d
                 + 4 * L * d * ff
                 + 8 * Q * d * d + 4 * Q * Q * d
                 + 4 * Q * d * d + 4 * L * d * d + 4 * Q * L * d
                 + 4 * Q * d * ff
                 + 2 * Q * (2 * d + 2 * d * d + 4 * d))
    return float(2 * L * (D * d + d * d) + 2 * (cfg.input_skch_dim * d + d * d) + cfg.num_layers * per_layer)
