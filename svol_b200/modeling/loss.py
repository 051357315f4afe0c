"""SetCriterion, drop-in for lib/modeling/loss.py: matching + losses of all decoder layers in three
kernel launches (svol_match, svol_match_localize, svol_criterion)."""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from .. import _lib
from .matcher import build_matcher, run_match, _check_status, _to_index_list


def _criterion_args(logits, boxes, flat, pred_idx, tgt_idx, losses, eos_coef):
    a = _lib.CriterionArgs()
    a.logits, a.boxes, a.tgt_boxes = logits.data_ptr(), boxes.data_ptr(), flat.tgt_boxes.data_ptr()
    a.pred_idx, a.tgt_idx = pred_idx.data_ptr(), tgt_idx.data_ptr()
    a.match_video, a.video_tgt_off = flat.match_video.data_ptr(), flat.video_tgt_off.data_ptr()
    a.losses = losses.data_ptr() if losses is not None else None
    a.NL, a.B, a.Q, a.K = logits.shape[0], logits.shape[1], logits.shape[2], flat.K
    a.eos_coef = float(eos_coef)
    return a


class _CriterionFn(torch.autograd.Function):
    """losses [NL,4] = (loss_label, class_error, loss_bbox, loss_giou); differentiable in logits / boxes."""

    @staticmethod
    def forward(ctx, logits, boxes, flat, pred_idx, tgt_idx, eos_coef):
        losses = torch.empty((logits.shape[0], 4), device=logits.device, dtype=torch.float32)
        a = _criterion_args(logits, boxes, flat, pred_idx, tgt_idx, losses, eos_coef)
        _lib.check(_lib.get_lib().svol_criterion(C.byref(a), _lib.stream_ptr()), "criterion")
        ctx.save_for_backward(logits, boxes, pred_idx, tgt_idx)
        ctx.flat, ctx.eos_coef = flat, eos_coef
        return losses

    @staticmethod
    def backward(ctx, grad):
        logits, boxes, pred_idx, tgt_idx = ctx.saved_tensors
        grad_w = grad[:, [0, 2, 3]].contiguous().float()           # class_error carries no gradient
        g_logits, g_boxes = torch.empty_like(logits), torch.empty_like(boxes)
        a = _criterion_args(logits, boxes, ctx.flat, pred_idx, tgt_idx, None, ctx.eos_coef)
        _lib.check(_lib.get_lib().svol_criterion_backward(C.byref(a), grad_w.data_ptr(), g_logits.data_ptr(),
                                                          g_boxes.data_ptr(), _lib.stream_ptr()), "criterion_backward")
        return g_logits, g_boxes, None, None, None, None


def _stack_layers(outputs):
    """[NL,B,Q,*] tensors in decoder-layer order (aux 0..NL-2, then the final layer).  Zero-copy when the
    dict came from ``svol_b200``'s SVANet (all entries are views of one buffer)."""
    lg, bx = outputs["pred_logits"], outputs["pred_boxes"]
    aux = outputs.get("aux_outputs", [])
    n = len(aux) + 1
    base_l, base_b = getattr(lg, "_base", None), getattr(bx, "_base", None)
    if (base_l is not None and base_b is not None and base_l.dim() == 4 and base_l.shape[0] == n
            and base_l.is_contiguous() and base_b.is_contiguous() and base_l.dtype == torch.float32
            and lg.data_ptr() == base_l[n - 1].data_ptr() and bx.data_ptr() == base_b[n - 1].data_ptr()
            and all(a["pred_logits"].data_ptr() == base_l[i].data_ptr() and a["pred_boxes"].data_ptr() == base_b[i].data_ptr()
                    for i, a in enumerate(aux))):
        return base_l, base_b
    logits = torch.stack([a["pred_logits"] for a in aux] + [lg]).float().contiguous()
    boxes = torch.stack([a["pred_boxes"] for a in aux] + [bx]).float().contiguous()
    return logits, boxes


class SetCriterion(nn.Module):
    """Same constructor, attributes (``weight_dict``, ``empty_weight``) and output keys as loss.py:10-157."""

    LOSS_NAMES = ("loss_label", "class_error", "loss_bbox", "loss_giou")

    def __init__(self, matcher, weight_dict, eos_coef, losses, bbox_type, sketch_head):
        super().__init__()
        if sketch_head != "svanet":
            raise NotImplementedError("only the live 'svanet' head is supported (sketch_detr is unreachable upstream)")
        unknown = set(losses) - {"labels", "boxes"}
        if unknown:
            raise NotImplementedError(f"losses {sorted(unknown)} are not produced by build_loss (loss.py:204)")
        self.matcher = matcher
        self.weight_dict = weight_dict
        self.losses = losses
        self.bbox_type = bbox_type
        self.sketch_head = sketch_head
        self.foreground_label, self.background_label = 0, 1
        self.eos_coef = eos_coef
        empty_weight = torch.ones(2)
        empty_weight[-1] = self.eos_coef
        self.register_buffer("empty_weight", empty_weight)
        self.last_indices = None          # (pred_idx, tgt_idx, flat) of the most recent call, on the device
        self.last_status = None

    def forward(self, outputs, targets):
        logits, boxes = _stack_layers(outputs)
        if not logits.is_cuda:
            raise RuntimeError("svol_b200 criterion needs CUDA tensors; there is no CPU fallback")
        m = self.matcher
        flat = m._flat(targets, logits.device, logits.shape[2])
        with torch.no_grad():
            pred_idx, tgt_idx, status, _ = run_match(logits.detach(), boxes.detach(), flat, m.cost_class, m.cost_bbox,
                                                     m.cost_giou)
        self.last_indices, self.last_status = (pred_idx, tgt_idx, flat), status
        L = _CriterionFn.apply(logits, boxes, flat, pred_idx, tgt_idx, float(self.eos_coef))
        n = logits.shape[0]
        want = {"labels": (0, 1), "boxes": (2, 3)}
        cols = [c for name in self.losses for c in want[name]]
        out = {}
        for li in [n - 1] + list(range(n - 1)):
            suffix = "" if li == n - 1 else f"_{li}"
            for c in cols:
                out[self.LOSS_NAMES[c] + suffix] = L[li, c]
        return out

    def check_status(self) -> None:
        """Raises like scipy would have (NaN / -inf cost entries).  Synchronises the device."""
        if self.last_status is not None:
            _check_status(self.last_status)

    def indices(self, layer: int = -1):
        """The matching of the most recent forward as the reference's list of CPU index tuples."""
        pred_idx, tgt_idx, flat = self.last_indices
        return _to_index_list(pred_idx[layer], tgt_idx[layer], flat)


def build_loss(args):
    """loss.py:192-213."""
    matcher = build_matcher(args)
    weight_dict = {"loss_bbox": args.set_cost_bbox, "loss_giou": args.set_cost_giou, "loss_label": args.set_cost_class}
    if args.aux_loss:
        aux = {}
        for i in range(args.num_layers - 1):
            aux.update({k + f"_{i}": v for k, v in weight_dict.items()})
        weight_dict.update(aux)
    return SetCriterion(matcher=matcher, weight_dict=weight_dict, eos_coef=args.eos_coef, losses=["labels", "boxes"],
                        bbox_type=args.bbox_type, sketch_head=args.sketch_head)
