"""CPU ORACLE -- test infrastructure, not product code.

A numpy restatement of the reference's SVOL hot path (head forward, Hungarian matching,
set losses).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module; the shipped package
``svol_b200`` never does (the CUDA extension is the only compute path there).

Pinning: every function below is checked in ``tests/test_oracle_golden.py`` against
golden vectors produced by importing the reference's own modules from /root/reference
(``tests/golden/make_golden.py``; torch 2.11.0 CPU, scipy 1.18.1), so parity is pinned to
the reference run in the build container.  Third-party arithmetic that the reference
delegates to (torch ``nn.MultiheadAttention`` / ``LayerNorm`` / ``F.gelu`` / ``cdist`` /
``cross_entropy``; ``scipy.optimize.linear_sum_assignment``) is restated from the published
definitions and pinned through those same vectors.

Each function cites the reference lines it follows.  ``dtype`` may be float32 (the
reference's working precision) or float64 (used to separate our error from the
reference's own fp32 rounding).
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np

try:                                    # exact erf for GELU; same image on the GPU box
    from scipy.special import erf as _erf
except Exception:                       # pragma: no cover
    _erf = np.vectorize(math.erf)

from . import lsap as _lsap


# --------------------------------------------------------------------------------------
# elementary layers (third-party torch ops, restated)
# --------------------------------------------------------------------------------------
def layer_norm(x, w, b, eps=1e-5):
    """torch.nn.LayerNorm over the last dim, biased variance, eps inside the sqrt."""
    mu = x.mean(-1, keepdims=True)
    xc = x - mu
    var = (xc * xc).mean(-1, keepdims=True)
    return xc / np.sqrt(var + x.dtype.type(eps)) * w + b


def linear(x, w, b):
    """torch.nn.Linear: x @ w.T + b."""
    return x @ w.T + b


def gelu(x):
    """F.gelu default (exact erf form); the activation of both transformer MLPs
    (cross_modal_transformer.py:11,163-179,185-194)."""
    return (0.5 * x * (1.0 + _erf(x / math.sqrt(2.0)))).astype(x.dtype)


def softmax(x, axis=-1):
    m = x.max(axis=axis, keepdims=True)
    e = np.exp(x - m)
    return e / e.sum(axis=axis, keepdims=True)


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def multihead_attention(q_in, k_in, v_in, in_w, in_b, out_w, out_b, nheads, key_padding_mask=None):
    """torch.nn.MultiheadAttention forward, batch-major: inputs (B, Lq|Lk, d).

    Returns (output (B,Lq,d), head-averaged weights (B,Lq,Lk)) exactly as the reference
    consumes them (cross_modal_transformer.py:124,139,147,154): packed in_proj, q scaled by
    1/sqrt(d_head) before QK^T, -inf on padded keys, softmax over keys, weights averaged
    over heads.
    """
    B, Lq, d = q_in.shape
    Lk = k_in.shape[1]
    dh = d // nheads
    q = linear(q_in, in_w[:d], in_b[:d]) * q_in.dtype.type(1.0 / math.sqrt(dh))
    k = linear(k_in, in_w[d:2 * d], in_b[d:2 * d])
    v = linear(v_in, in_w[2 * d:], in_b[2 * d:])
    q = q.reshape(B, Lq, nheads, dh).transpose(0, 2, 1, 3)
    k = k.reshape(B, Lk, nheads, dh).transpose(0, 2, 1, 3)
    v = v.reshape(B, Lk, nheads, dh).transpose(0, 2, 1, 3)
    s = q @ k.transpose(0, 1, 3, 2)                                  # (B,H,Lq,Lk)
    if key_padding_mask is not None:
        s = np.where(key_padding_mask[:, None, None, :], -np.inf, s).astype(q.dtype)
    p = softmax(s, axis=-1)
    o = (p @ v).transpose(0, 2, 1, 3).reshape(B, Lq, d)
    return linear(o, out_w, out_b), p.mean(axis=1)


def position_embedding_sine(mask, num_pos_feats, temperature=10000.0, dtype=np.float32):
    """PositionEmbeddingSine.forward with normalize=True (position_encoding.py:51-71).
    ``mask``: (B, L) bool, True = valid.  Returns (B, L, num_pos_feats)."""
    # The reference always builds this table in float32 (cumsum dtype and arange dtype are
    # hard-coded, :57,:63) whatever the model dtype is; so does the oracle.
    f = np.float32
    x = np.cumsum(mask.astype(f), axis=1, dtype=f)
    x = x / (x[:, -1:] + f(1e-6)) * f(2.0 * math.pi)
    i = np.arange(num_pos_feats, dtype=f)
    dim_t = np.power(f(temperature), (f(2.0) * np.trunc(i / f(2.0)) / f(num_pos_feats)).astype(f)).astype(f)
    pos = (x[:, :, None] / dim_t).astype(f)
    out = np.stack((np.sin(pos[:, :, 0::2]), np.cos(pos[:, :, 1::2])), axis=3)
    return out.reshape(mask.shape[0], mask.shape[1], num_pos_feats).astype(dtype)


# --------------------------------------------------------------------------------------
# head forward
# --------------------------------------------------------------------------------------
def input_projection(x, sd, prefix, n_input_proj):
    """nn.Sequential of LinearLayer: LN -> Dropout(eval: identity) -> Linear (-> ReLU)
    with ReLU on every layer except the last (svanet.py:49-60,159-181)."""
    for i in range(n_input_proj):
        x = layer_norm(x, sd[f"{prefix}.{i}.LayerNorm.weight"], sd[f"{prefix}.{i}.LayerNorm.bias"])
        x = linear(x, sd[f"{prefix}.{i}.net.1.weight"], sd[f"{prefix}.{i}.net.1.bias"])
        if i != n_input_proj - 1:
            x = np.maximum(x, 0)
    return x


def _mha_params(sd, prefix):
    return (sd[prefix + ".in_proj_weight"], sd[prefix + ".in_proj_bias"],
            sd[prefix + ".out_proj.weight"], sd[prefix + ".out_proj.bias"])


def _mlp(x, sd, prefix):
    h = gelu(linear(x, sd[prefix + ".fc1.weight"], sd[prefix + ".fc1.bias"]))
    return linear(h, sd[prefix + ".fc2.weight"], sd[prefix + ".fc2.bias"])


def transformer_layer(sd, p, nheads, vid, skch, out, vid_pad_mask, vid_pos, query_pos):
    """CrossModalTransformerLayer.forward (cross_modal_transformer.py:105-160), batch-major.
    Returns (mem, out, att1)."""
    ln = lambda x, n: layer_norm(x, sd[f"{p}.norm{n}.weight"], sd[f"{p}.norm{n}.bias"])
    # (a) sketch -> video attention weights gate the video tokens  (:122-127)
    kv = vid + vid_pos
    _, att1 = multihead_attention(skch, kv, kv, *_mha_params(sd, p + ".sketch_video_cross_attn"), nheads)
    mem = ln(vid + att1.transpose(0, 2, 1) * vid, 1)
    # (b) video self-attention (no padding mask) + FFN  (:137-143)
    qk = mem + vid_pos
    a, _ = multihead_attention(qk, qk, mem, *_mha_params(sd, p + ".content_self_attn"), nheads)
    mem = ln(a + mem, 2)
    mem = ln(mem + _mlp(mem, sd, p + ".mlp1"), 3)
    # (c) query self-attention  (:145-149)
    qk = out + query_pos
    a, _ = multihead_attention(qk, qk, out, *_mha_params(sd, p + ".token_self_attn"), nheads)
    out = ln(a + out, 4)
    # (d) query -> video cross-attention with key padding mask + FFN  (:151-158)
    a, _ = multihead_attention(out + query_pos, mem + vid_pos, mem,
                               *_mha_params(sd, p + ".content_token_cross_attn"), nheads,
                               key_padding_mask=vid_pad_mask)
    out = ln(out + a, 5)
    out = ln(out + _mlp(out, sd, p + ".mlp2"), 6)
    return mem, out, att1


def svanet_forward(sd: Dict[str, np.ndarray], src_sketch, src_sketch_mask, src_video, src_video_mask,
                   nheads=8, n_input_proj=2, dtype=np.float32, return_intermediates=False):
    """SVANet.forward in eval mode (svanet.py:65-141) on top of CrossModalTransformer.forward
    (cross_modal_transformer.py:27-81).  Returns the reference's output dict with numpy leaves
    (plus 'hs' and per-layer 'mem' when ``return_intermediates``)."""
    sd = {k: np.asarray(v, dtype=dtype) for k, v in sd.items()}
    vid = input_projection(np.asarray(src_video, dtype), sd, "input_video_proj", n_input_proj)
    mask_vid = np.asarray(src_video_mask) != 0
    d = vid.shape[-1]
    pos_vid = position_embedding_sine(mask_vid, d, dtype=dtype)
    skch = input_projection(np.asarray(src_sketch, dtype), sd, "input_sketch_proj", n_input_proj)
    B = vid.shape[0]
    query_pos = np.broadcast_to(sd["query_embed.weight"][None], (B,) + sd["query_embed.weight"].shape)
    out = np.zeros_like(query_pos)
    num_layers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("transformer.layers."))
    hs, mems, atts = [], [], []
    mem = vid
    for li in range(num_layers):
        mem, out, att1 = transformer_layer(sd, f"transformer.layers.{li}", nheads, mem, skch, out,
                                           ~mask_vid, pos_vid, query_pos)
        hs.append(out); mems.append(mem); atts.append(att1)
    hs = np.stack(hs)                                                   # (N,B,Q,d)
    logits = linear(hs, sd["class_embed.weight"], sd["class_embed.bias"])
    x = hs
    for i in range(3):                                                  # MLP(d,d,4,3)  svanet.py:144-156
        x = linear(x, sd[f"bbox_embed.layers.{i}.weight"], sd[f"bbox_embed.layers.{i}.bias"])
        if i < 2:
            x = np.maximum(x, 0)
    boxes = sigmoid(x)
    res = {"pred_logits": logits[-1], "pred_boxes": boxes[-1],
           "aux_outputs": [{"pred_logits": a, "pred_boxes": b} for a, b in zip(logits[:-1], boxes[:-1])]}
    if return_intermediates:
        res.update(hs=hs, mem=np.stack(mems), att1=np.stack(atts), all_logits=logits, all_boxes=boxes,
                   src_vid=vid, pos_vid=pos_vid, src_skch=skch)
    return res


# --------------------------------------------------------------------------------------
# box utilities (lib/utils/box_utils.py)
# --------------------------------------------------------------------------------------
def box_cxcywh_to_xyxy(x):
    """box_utils.py:9-13"""
    h = x.dtype.type(0.5)
    cx, cy, w, hh = x[..., 0], x[..., 1], x[..., 2], x[..., 3]
    return np.stack([cx - h * w, cy - h * hh, cx + h * w, cy + h * hh], axis=-1)


def box_iou(b1, b2):
    """box_utils.py:24-37 (pairwise)"""
    a1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    lt = np.maximum(b1[:, None, :2], b2[None, :, :2])
    rb = np.minimum(b1[:, None, 2:], b2[None, :, 2:])
    wh = np.clip(rb - lt, 0, None)
    inter = wh[..., 0] * wh[..., 1]
    union = a1[:, None] + a2[None, :] - inter
    return inter / union, union


def generalized_box_iou(b1, b2):
    """box_utils.py:40-61 (pairwise, with the reference's degenerate-box asserts)."""
    assert (b1[:, 2:] >= b1[:, :2]).all()
    assert (b2[:, 2:] >= b2[:, :2]).all()
    iou, union = box_iou(b1, b2)
    lt = np.minimum(b1[:, None, :2], b2[None, :, :2])
    rb = np.maximum(b1[:, None, 2:], b2[None, :, 2:])
    wh = np.clip(rb - lt, 0, None)
    area = wh[..., 0] * wh[..., 1]
    return iou - (area - union) / area


# --------------------------------------------------------------------------------------
# targets flattening, matcher
# --------------------------------------------------------------------------------------
def flatten_targets(targets: Sequence[dict]):
    """The walk of matcher.py:62-70 / loss.py:78-85: video -> frame (dict order) -> instance.
    Returns (boxes (sum,4) float32, num_boxes_per_frame list of B*T ints, boxes per video list)."""
    boxes, num_boxes, per_video = [], [], []
    for t in targets:
        num_boxes.extend(int(n) for n in t["num_boxes_per_frame"])
        cnt = 0
        for frame in t["bboxes"].values():
            for inst in frame:
                boxes.append(np.asarray(inst["bbox"], dtype=np.float32))
                cnt += 1
        per_video.append(cnt)
    return np.stack(boxes).astype(np.float32), num_boxes, per_video


def cost_matrix(logits, boxes, tgt_boxes, w_class, w_bbox, w_giou):
    """matcher.py:59-85 for one set of predictions: logits (n,2), boxes (n,4) vs tgt (m,4).
    fp32, the reference's operation order: C = w_bbox*L1 + w_giou*(-GIoU) + w_class*(-p_fg)."""
    f = np.float32
    prob = softmax(logits.astype(f), axis=-1)
    cost_class = -prob[:, 0:1]                                           # foreground label 0  (:71-76)
    cost_bbox = np.abs(boxes[:, None, :].astype(f) - tgt_boxes[None, :, :].astype(f)).sum(-1, dtype=f)   # cdist p=1 (:79)
    cost_giou = -generalized_box_iou(box_cxcywh_to_xyxy(boxes.astype(f)), box_cxcywh_to_xyxy(tgt_boxes.astype(f)))
    return (f(w_bbox) * cost_bbox + f(w_giou) * cost_giou + f(w_class) * cost_class).astype(f)


def per_frame_matcher(logits, boxes, targets, num_frames, q_per_frame,
                      w_class=2.0, w_bbox=5.0, w_giou=1.0, solver=None):
    """PerFrameMatcher.forward (matcher.py:38-119).  logits (B,Q,2), boxes (B,Q,4).
    Only the block-diagonal cost entries the reference reads (:92-93) are formed.
    Returns [(pred_idx int64, tgt_idx int64)] per video, in the reference's order and with
    its 'subtract the minimum matched target index' localisation (:114-115)."""
    solver = solver or _lsap.linear_sum_assignment
    B, Q = boxes.shape[:2]
    assert Q == num_frames * q_per_frame                                 # matcher.py:56
    tgt, num_boxes, _ = flatten_targets(targets)
    offs = np.concatenate([[0], np.cumsum(num_boxes)]).astype(np.int64)
    result = []
    for b in range(B):
        pred_v, tgt_v = [], []
        for t in range(num_frames):
            i = b * num_frames + t
            n = num_boxes[i]
            rows = slice(t * q_per_frame, (t + 1) * q_per_frame)
            if n == 0:
                continue
            C = cost_matrix(logits[b, rows], boxes[b, rows], tgt[offs[i]:offs[i] + n], w_class, w_bbox, w_giou)
            r, c = solver(C)
            pred_v.extend((r + t * q_per_frame).tolist())
            tgt_v.extend((c + offs[i]).tolist())
        tgt_v = np.asarray(tgt_v, dtype=np.int64)
        tgt_v = tgt_v - tgt_v.min()                                       # matcher.py:114-115
        result.append((np.asarray(pred_v, dtype=np.int64), tgt_v))
    return result


def video_matcher(logits, boxes, targets, w_class=2.0, w_bbox=5.0, w_giou=1.0, solver=None):
    """HungarianMatcher.forward (matcher.py:131-159): one Q x n_v assignment per video."""
    solver = solver or _lsap.linear_sum_assignment
    tgt, _, per_video = flatten_targets(targets)
    offs = np.concatenate([[0], np.cumsum(per_video)]).astype(np.int64)
    out = []
    for b in range(boxes.shape[0]):
        C = cost_matrix(logits[b], boxes[b], tgt[offs[b]:offs[b + 1]], w_class, w_bbox, w_giou)
        r, c = solver(C)
        out.append((np.asarray(r, np.int64), np.asarray(c, np.int64)))
    return out


# --------------------------------------------------------------------------------------
# criterion
# --------------------------------------------------------------------------------------
def loss_labels(logits, indices, eos_coef=0.1, dtype=np.float32):
    """SetCriterion.loss_labels (loss.py:39-60): weighted CE, mean over all B*Q entries
    (not divided by the weight sum), plus class_error = 100 - top-1 accuracy on the matched
    queries (model_utils.py:4-21)."""
    x = logits.astype(dtype)
    B, Q, _ = x.shape
    target = np.ones((B, Q), np.int64)                                   # background
    for b, (src, _) in enumerate(indices):
        target[b, src] = 0                                               # foreground
    m = x.max(-1, keepdims=True)
    lse = (m + np.log(np.exp(x - m).sum(-1, keepdims=True)))[..., 0]
    nll = lse - np.take_along_axis(x, target[..., None], -1)[..., 0]
    w = np.where(target == 0, dtype(1.0), dtype(eos_coef))
    loss = (w * nll).mean(dtype=dtype)
    matched = np.concatenate([x[b, src] for b, (src, _) in enumerate(indices)])
    # topk(1) picks index 0 unless logit[1] is strictly larger
    correct = (matched[:, 1] <= matched[:, 0]).sum()
    class_error = dtype(100.0) - dtype(correct) * dtype(100.0 / matched.shape[0])
    return {"loss_label": loss, "class_error": class_error}


def loss_boxes(boxes, targets, indices, dtype=np.float32):
    """SetCriterion.loss_boxes (loss.py:76-103): L1 mean over K*4 elements and
    mean(1 - GIoU) over the K matched pairs."""
    _, _, per_video = flatten_targets(targets)
    tgt_all, _, _ = flatten_targets(targets)
    offs = np.concatenate([[0], np.cumsum(per_video)]).astype(np.int64)
    src = np.concatenate([boxes[b, s] for b, (s, _) in enumerate(indices)]).astype(dtype)
    tgt = np.concatenate([tgt_all[offs[b] + t] for b, (_, t) in enumerate(indices)]).astype(dtype)
    l1 = np.abs(src - tgt).mean(dtype=dtype)
    sx, tx = box_cxcywh_to_xyxy(src), box_cxcywh_to_xyxy(tgt)
    # diagonal of the pairwise matrix == elementwise GIoU of the pairs
    a1 = (sx[:, 2] - sx[:, 0]) * (sx[:, 3] - sx[:, 1])
    a2 = (tx[:, 2] - tx[:, 0]) * (tx[:, 3] - tx[:, 1])
    wh = np.clip(np.minimum(sx[:, 2:], tx[:, 2:]) - np.maximum(sx[:, :2], tx[:, :2]), 0, None)
    inter = wh[:, 0] * wh[:, 1]
    union = a1 + a2 - inter
    iou = inter / union
    whc = np.clip(np.maximum(sx[:, 2:], tx[:, 2:]) - np.minimum(sx[:, :2], tx[:, :2]), 0, None)
    area = whc[:, 0] * whc[:, 1]
    giou = iou - (area - union) / area
    return {"loss_bbox": l1, "loss_giou": (1 - giou).mean(dtype=dtype)}


def set_criterion(outputs, targets, cfg, dtype=np.float32, solver=None, return_indices=False):
    """SetCriterion.forward (loss.py:126-157): matcher + both losses for the last layer and
    for every aux layer (keys suffixed ``_i``)."""
    def match(lg, bx):
        if cfg.matcher == "per_frame_matcher":
            return per_frame_matcher(lg, bx, targets, cfg.num_frames, cfg.num_queries_per_frame,
                                     cfg.set_cost_class, cfg.set_cost_bbox, cfg.set_cost_giou, solver)
        return video_matcher(lg, bx, targets, cfg.set_cost_class, cfg.set_cost_bbox, cfg.set_cost_giou, solver)

    losses, all_idx = {}, []
    layers = [(outputs["pred_logits"], outputs["pred_boxes"], "")]
    for i, aux in enumerate(outputs.get("aux_outputs", [])):
        layers.append((aux["pred_logits"], aux["pred_boxes"], f"_{i}"))
    for lg, bx, suffix in layers:
        idx = match(np.asarray(lg, np.float32), np.asarray(bx, np.float32))
        all_idx.append(idx)
        for k, v in loss_labels(np.asarray(lg), idx, cfg.eos_coef, dtype).items():
            losses[k + suffix] = v
        for k, v in loss_boxes(np.asarray(bx), targets, idx, dtype).items():
            losses[k + suffix] = v
    return (losses, all_idx) if return_indices else losses


def weight_dict(cfg) -> Dict[str, float]:
    """build_loss (loss.py:192-213)."""
    base = {"loss_bbox": cfg.set_cost_bbox, "loss_giou": cfg.set_cost_giou, "loss_label": cfg.set_cost_class}
    wd = dict(base)
    if cfg.aux_loss:
        for i in range(cfg.num_layers - 1):
            wd.update({f"{k}_{i}": v for k, v in base.items()})
    return wd


# --------------------------------------------------------------------------------------
# inference post-processing (test.py:133-158) -- SURVEY 8(f)-1
# --------------------------------------------------------------------------------------
def postprocess(logits, boxes, num_frames):
    """softmax foreground score, clamp(xyxy, 0, 1), per-frame chunks sorted by score
    (descending, stable like Python's ``sorted``).  Returns (B, T, q_f, 5) float32 and the
    per-frame permutation (B, T, q_f) int64."""
    prob = softmax(logits.astype(np.float32), -1)[..., 0]
    xyxy = np.clip(box_cxcywh_to_xyxy(boxes.astype(np.float32)), 0, 1)
    B, Q = prob.shape
    qf = Q // num_frames
    preds = np.concatenate([xyxy, prob[..., None]], -1).reshape(B, num_frames, qf, 5)
    order = np.argsort(-preds[..., 4], axis=-1, kind="stable")
    return np.take_along_axis(preds, order[..., None], axis=2), order
