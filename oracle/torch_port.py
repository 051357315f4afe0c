"""CPU ORACLE (second restatement) -- test infrastructure, not product code.

The reference's hot path restated with THE SAME LIBRARY CALLS the reference makes on a CPU: PyTorch ATen ops
(``F.multi_head_attention_forward`` with ``need_weights=True`` exactly as ``nn.MultiheadAttention`` runs it,
``F.layer_norm``, ``F.gelu``, ``torch.cdist``, ``F.cross_entropy``) and ``scipy.optimize.linear_sum_assignment``,
including the reference's inefficiencies (the full cross-batch cost matrix, one scipy call per frame, the K x K
GIoU whose diagonal is the loss).  It exists because ``bench.py --impl reference`` / ``cpu_baseline`` should time
what the reference's CPU path actually costs: the numpy restatement in ``svol_oracle.py`` is the parity checker
(fp64-capable, block-diagonal costs) but is ~2.5x slower than ATen on the same cores and skips work the reference
does.  Pinned against the same golden vectors (tests/test_oracle_golden.py).  Only tests/ and bench.py import it.

Each function cites the reference lines it follows; nothing here is copied from /root/reference.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import numpy as np
import torch
import torch.nn.functional as F

try:
    from scipy.optimize import linear_sum_assignment as _scipy_lsap
except Exception:          # pragma: no cover -- the image ships scipy; the C restatement is the fallback
    _scipy_lsap = None


def _lsap(cost: torch.Tensor):
    if _scipy_lsap is not None:
        return _scipy_lsap(cost)
    from . import lsap
    return lsap.linear_sum_assignment(cost.numpy())


# ------------------------------------------------------------------------------------------ head forward
def dropout_mask(rows: int, cols: int, p: float, seed: int, site: int) -> np.ndarray:
    """Host restatement of the CUDA path's counter-based dropout mask (include/svol_b200.h, common.cuh: dropout_keep):
    element idx = row * cols + col of site `site` is kept iff the top 32 bits of
    splitmix64(idx + (8 * seed + site) * 0x9E3779B97F4A7C15) are >= p * 2^32.  Returns the fp32 multiplier
    keep / (1 - p), shape (rows, cols).  (nn.Dropout's own Philox stream cannot be reproduced by another
    implementation; parity under dropout is checked with THIS mask on both sides.)"""
    m64 = (1 << 64) - 1
    key = np.uint64((((8 * seed + site) & m64) * 0x9E3779B97F4A7C15) & m64)
    with np.errstate(over="ignore"):
        z = np.arange(rows * cols, dtype=np.uint64) + key
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    thr = np.uint64(min(int(np.float32(p) * np.float32(4294967296.0)), 4294967040))
    keep = (z >> np.uint64(32)) >= thr
    scale = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    return (keep.astype(np.float32) * scale).reshape(rows, cols)


def _input_proj(x, sd, prefix, n, dropout=None, first_site=0):
    """LinearLayer x n: LayerNorm -> Dropout (eval: identity) -> Linear -> ReLU except after the last
    (svanet.py:49-60,159-181).  ``dropout`` = (p, seed): train mode with the CUDA path's mask at sites first_site + i."""
    for i in range(n):
        x = F.layer_norm(x, x.shape[-1:], sd[f"{prefix}.{i}.LayerNorm.weight"], sd[f"{prefix}.{i}.LayerNorm.bias"])
        if dropout is not None and dropout[0] > 0:
            rows = x.numel() // x.shape[-1]
            mask = torch.from_numpy(dropout_mask(rows, x.shape[-1], dropout[0], dropout[1], first_site + i)).to(x.dtype)
            x = x * mask.reshape(x.shape)
        x = F.linear(x, sd[f"{prefix}.{i}.net.1.weight"], sd[f"{prefix}.{i}.net.1.bias"])
        if i != n - 1:
            x = F.relu(x)
    return x


def _pos_sine(mask, d, temperature=10000.0):
    """PositionEmbeddingSine(normalize=True) (position_encoding.py:51-71); mask (B,L) bool, True = valid."""
    x = mask.cumsum(1, dtype=torch.float32)
    x = x / (x[:, -1:] + 1e-6) * (2 * math.pi)
    i = torch.arange(d, dtype=torch.float32, device=mask.device)
    dim_t = temperature ** (2 * torch.div(i, 2, rounding_mode="trunc") / d)
    p = x[:, :, None] / dim_t
    return torch.stack((p[:, :, 0::2].sin(), p[:, :, 1::2].cos()), dim=3).flatten(2)


def _mha(q, k, v, sd, prefix, nheads, key_padding_mask=None):
    """nn.MultiheadAttention(d, nheads).forward on (S,B,d) tensors, dropout 0, need_weights=True (the default the
    reference uses at cross_modal_transformer.py:124,139,147,154 -> scores are materialised and head-averaged)."""
    return F.multi_head_attention_forward(
        q, k, v, q.shape[-1], nheads, sd[prefix + ".in_proj_weight"], sd[prefix + ".in_proj_bias"], None, None, False,
        0.0, sd[prefix + ".out_proj.weight"], sd[prefix + ".out_proj.bias"], training=False,
        key_padding_mask=key_padding_mask, need_weights=True)


def _ffn(x, sd, prefix):
    return F.linear(F.gelu(F.linear(x, sd[prefix + ".fc1.weight"], sd[prefix + ".fc1.bias"])),
                    sd[prefix + ".fc2.weight"], sd[prefix + ".fc2.bias"])


@torch.no_grad()
def svanet_forward(sd: Dict[str, torch.Tensor], src_sketch, src_sketch_mask, src_video, src_video_mask, nheads=8,
                   n_input_proj=2, dropout=None):
    """SVANet.forward (svanet.py:65-141) + CrossModalTransformer.forward (cross_modal_transformer.py:27-81,105-160),
    eval mode, fp32, sequence-major (S,B,d) inside the transformer like the reference.  ``dropout`` = (p, seed): the
    train-mode input dropout with the CUDA path's mask (sites 0,1 frame tokens; 2,3 sketch)."""
    vid = _input_proj(src_video, sd, "input_video_proj", n_input_proj, dropout, 0)
    skch = _input_proj(src_sketch, sd, "input_sketch_proj", n_input_proj, dropout, 2)
    mask_vid = src_video_mask != 0
    pos = _pos_sine(mask_vid, vid.shape[-1]).permute(1, 0, 2)
    pad = ~mask_vid
    B = vid.shape[0]
    mem, skch = vid.permute(1, 0, 2), skch.permute(1, 0, 2)
    qpos = sd["query_embed.weight"][:, None, :].repeat(1, B, 1)
    out = torch.zeros_like(qpos)
    n_layers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("transformer.layers."))
    hs = []
    for li in range(n_layers):
        p = f"transformer.layers.{li}"
        ln = lambda x, n: F.layer_norm(x, x.shape[-1:], sd[f"{p}.norm{n}.weight"], sd[f"{p}.norm{n}.bias"])
        kv = mem + pos
        _, att1 = _mha(skch, kv, kv, sd, p + ".sketch_video_cross_attn", nheads)                  # :122-127
        mem = ln(mem + att1.permute(2, 0, 1) * mem, 1)
        qk = mem + pos
        mem = ln(_mha(qk, qk, mem, sd, p + ".content_self_attn", nheads)[0] + mem, 2)             # :137-141
        mem = ln(mem + _ffn(mem, sd, p + ".mlp1"), 3)                                             # :142-143
        qk = out + qpos
        out = ln(_mha(qk, qk, out, sd, p + ".token_self_attn", nheads)[0] + out, 4)               # :145-149
        out = ln(out + _mha(out + qpos, mem + pos, mem, sd, p + ".content_token_cross_attn", nheads, pad)[0], 5)
        out = ln(out + _ffn(out, sd, p + ".mlp2"), 6)                                             # :151-158
        hs.append(out.transpose(0, 1))
    hs = torch.stack(hs)
    logits = F.linear(hs, sd["class_embed.weight"], sd["class_embed.bias"])                       # svanet.py:125
    x = hs
    for i in range(3):                                                                            # svanet.py:144-156
        x = F.linear(x, sd[f"bbox_embed.layers.{i}.weight"], sd[f"bbox_embed.layers.{i}.bias"])
        if i < 2:
            x = F.relu(x)
    boxes = x.sigmoid()
    return {"pred_logits": logits[-1], "pred_boxes": boxes[-1],
            "aux_outputs": [{"pred_logits": a, "pred_boxes": b} for a, b in zip(logits[:-1], boxes[:-1])]}


# ------------------------------------------------------------------------------------------ boxes / matcher / losses
def _xyxy(b):                                                                                     # box_utils.py:9-13
    cx, cy, w, h = b.unbind(-1)
    return torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], dim=-1)


def _giou(a, b):
    """generalized_box_iou on xyxy boxes, full pairwise (N,M) (box_utils.py:24-61)."""
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    wh = (torch.min(a[:, None, 2:], b[:, 2:]) - torch.max(a[:, None, :2], b[:, :2])).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    union = area_a[:, None] + area_b - inter
    iou = inter / union
    whc = (torch.max(a[:, None, 2:], b[:, 2:]) - torch.min(a[:, None, :2], b[:, :2])).clamp(min=0)
    hull = whc[..., 0] * whc[..., 1]
    return iou - (hull - union) / hull


def _target_boxes(targets):
    return torch.stack([torch.as_tensor(np.asarray(o["bbox"]), dtype=torch.float32)
                        for t in targets for fr in t["bboxes"].values() for o in fr])


@torch.no_grad()
def per_frame_matcher(logits, boxes, targets, num_frames, q_per_frame, w_class=2.0, w_bbox=5.0, w_giou=1.0):
    """PerFrameMatcher.forward (matcher.py:38-119) the way the reference computes it: softmax, the FULL cross-batch
    (B*Q) x sum(n) cost matrix, then one scipy call per frame on its block, then regrouping per video."""
    B, Q = logits.shape[:2]
    prob = logits.flatten(0, 1).softmax(-1)
    obox = boxes.flatten(0, 1)
    tgt = _target_boxes(targets).to(obox.device)                                                  # matcher.py:70
    num_boxes = [int(n) for t in targets for n in t["num_boxes_per_frame"]]
    cost = w_bbox * torch.cdist(obox, tgt, p=1) - w_giou * _giou(_xyxy(obox), _xyxy(tgt)) - w_class * prob[:, :1]
    cost = cost.view(B * num_frames, q_per_frame, -1).cpu()                                       # matcher.py:86
    res, off = [], 0
    for b in range(B):
        pv, tv = [], []
        for t in range(num_frames):
            i = b * num_frames + t
            n = num_boxes[i]
            if n:
                r, c = _lsap(cost[i, :, off:off + n])
                pv.append(torch.as_tensor(r, dtype=torch.int64) + t * q_per_frame)
                tv.append(torch.as_tensor(c, dtype=torch.int64) + off)
            off += n
        tv = torch.cat(tv)
        res.append((torch.cat(pv), tv - tv.min()))                                                # matcher.py:114-115
    return res


@torch.no_grad()
def video_matcher(logits, boxes, targets, w_class=2.0, w_bbox=5.0, w_giou=1.0):
    """HungarianMatcher.forward (matcher.py:131-159)."""
    B, Q = logits.shape[:2]
    prob = logits.flatten(0, 1).softmax(-1)
    obox = boxes.flatten(0, 1)
    tgt = _target_boxes(targets).to(obox.device)
    sizes = [sum(len(fr) for fr in t["bboxes"].values()) for t in targets]
    cost = w_bbox * torch.cdist(obox, tgt, p=1) - w_giou * _giou(_xyxy(obox), _xyxy(tgt)) - w_class * prob[:, :1]
    cost = cost.view(B, Q, -1).cpu()                                                              # matcher.py:152
    res = []
    for b, c in enumerate(cost.split(sizes, -1)):
        r, cc = _lsap(c[b])
        res.append((torch.as_tensor(r, dtype=torch.int64), torch.as_tensor(cc, dtype=torch.int64)))
    return res


def set_criterion(outputs, targets, cfg):
    """SetCriterion.forward (loss.py:126-157) with loss_labels (:39-60) and loss_boxes (:76-103)."""
    def match(lg, bx):
        if cfg.matcher == "per_frame_matcher":
            return per_frame_matcher(lg, bx, targets, cfg.num_frames, cfg.num_queries_per_frame, cfg.set_cost_class,
                                     cfg.set_cost_bbox, cfg.set_cost_giou)
        return video_matcher(lg, bx, targets, cfg.set_cost_class, cfg.set_cost_bbox, cfg.set_cost_giou)

    dev = outputs["pred_logits"].device
    weight = torch.tensor([1.0, cfg.eos_coef], device=dev)
    per_video = [torch.stack([torch.as_tensor(np.asarray(o["bbox"]), dtype=torch.float32)
                              for fr in t["bboxes"].values() for o in fr]).to(dev) for t in targets]
    losses, all_idx = {}, []
    layers = [(outputs["pred_logits"], outputs["pred_boxes"], "")]
    layers += [(a["pred_logits"], a["pred_boxes"], f"_{i}") for i, a in enumerate(outputs.get("aux_outputs", []))]
    for lg, bx, suffix in layers:
        idx = match(lg, bx)
        all_idx.append(idx)
        bidx = torch.cat([torch.full_like(s, i) for i, (s, _) in enumerate(idx)])         # CPU int64, like the reference
        sidx = torch.cat([s for s, _ in idx])
        tc = torch.full(lg.shape[:2], 1, dtype=torch.int64, device=dev)
        tc[bidx, sidx] = 0
        losses["loss_label" + suffix] = F.cross_entropy(lg.transpose(1, 2), tc, weight, reduction="none").mean()
        matched = lg[bidx, sidx]
        acc = (matched.topk(1, 1)[1][:, 0] == 0).float().sum() * (100.0 / matched.shape[0])       # model_utils.py:4-21
        losses["class_error" + suffix] = 100 - acc
        src = bx[bidx, sidx]
        tgt = torch.cat([pv[t] for pv, (_, t) in zip(per_video, idx)])
        losses["loss_bbox" + suffix] = F.l1_loss(src, tgt, reduction="none").mean()
        losses["loss_giou" + suffix] = (1 - torch.diag(_giou(_xyxy(src), _xyxy(tgt)))).mean()     # K x K, diagonal used
    return losses, all_idx


# ------------------------------------------------------------------------------------------ training step (autograd)
def _stack_outputs(out):
    logits = torch.stack([a["pred_logits"] for a in out.get("aux_outputs", [])] + [out["pred_logits"]])
    boxes = torch.stack([a["pred_boxes"] for a in out.get("aux_outputs", [])] + [out["pred_boxes"]])
    return logits, boxes


def head_gradients(sd: Dict[str, torch.Tensor], src_sketch, src_sketch_mask, src_video, src_video_mask, grad_logits,
                   grad_boxes, nheads=8, n_input_proj=2, dtype=torch.float32, dropout=None):
    """Parameter gradients of the head for given upstream gradients w.r.t. the stacked (logits [NL,B,Q,2], boxes
    [NL,B,Q,4]) of all decoder layers: what ``loss.backward()`` (train.py:229) propagates through
    ``SVANet.forward`` in train mode with the input dropout at 0.  Returns ({name: grad}, logits, boxes)."""
    names = [k for k, v in sd.items() if v.is_floating_point()]
    leaf = {k: (sd[k].detach().to(dtype).clone().requires_grad_(True) if k in names else sd[k]) for k in sd}
    c = lambda t: torch.as_tensor(t).to(dtype)
    with torch.enable_grad():
        out = svanet_forward.__wrapped__(leaf, c(src_sketch), c(src_sketch_mask), c(src_video), c(src_video_mask), nheads,
                                         n_input_proj, dropout)
        logits, boxes = _stack_outputs(out)
        grads = torch.autograd.grad([logits, boxes], [leaf[k] for k in names], [c(grad_logits), c(grad_boxes)],
                                    allow_unused=True)
    return {k: g for k, g in zip(names, grads) if g is not None}, logits.detach(), boxes.detach()


def training_step_gradients(sd: Dict[str, torch.Tensor], src_sketch, src_sketch_mask, src_video, src_video_mask, targets,
                            cfg, weight_dict: Dict[str, float], dtype=torch.float32, dropout=None):
    """train.py:222-229: outputs = model(...); loss_dict = criterion(outputs, targets); losses = sum(loss_dict[k] *
    weight_dict[k]); losses.backward().  Returns ({name: grad}, {loss name: value}, matching indices)."""
    names = [k for k, v in sd.items() if v.is_floating_point()]
    leaf = {k: (sd[k].detach().to(dtype).clone().requires_grad_(True) if k in names else sd[k]) for k in sd}
    c = lambda t: torch.as_tensor(t).to(dtype)
    with torch.enable_grad():
        out = svanet_forward.__wrapped__(leaf, c(src_sketch), c(src_sketch_mask), c(src_video), c(src_video_mask), cfg.nheads,
                                         cfg.n_input_proj, dropout)
        out32 = {"pred_logits": out["pred_logits"].float(), "pred_boxes": out["pred_boxes"].float(),
                 "aux_outputs": [{k: v.float() for k, v in a.items()} for a in out["aux_outputs"]]}
        losses, idx = set_criterion(out32, targets, cfg)
        total = sum(losses[k] * weight_dict[k] for k in losses if k in weight_dict)
        grads = torch.autograd.grad(total, [leaf[k] for k in names], allow_unused=True)
    return ({k: g for k, g in zip(names, grads) if g is not None}, {k: float(v.detach()) for k, v in losses.items()}, idx)


def state_dict_to_torch(sd: Dict[str, np.ndarray]) -> Dict[str, torch.Tensor]:
    return {k: torch.from_numpy(np.ascontiguousarray(v)).float() for k, v in sd.items()}
