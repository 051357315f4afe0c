"""SetCriterion, drop-in for lib/modeling/loss.py: matching + losses of all decoder layers in three kernel launches
(svol_match = cost blocks + assignment, its finalize kernel, svol_criterion) that replay as ONE CUDA graph.

The launches work on a static workspace (``_Workspace``): the batch's packed targets are copied into a device buffer
whose arrays sit at fixed addresses, the sizes that change from batch to batch (number of boxes, of matched pairs) are
read by the kernels from that buffer, and the index / status / loss outputs are plan-owned.  So a step allocates
nothing, fills no buffer from the host (the matcher's status word is published and cleared by the call's own finalize
kernel) and -- when the predictions are the head's own static output buffers -- costs one graph launch.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from .. import _lib
from .matcher import COST_SMEM_LIMIT, build_matcher, fill_match_args, _check_status, _to_index_list
from .targets import static_views


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class _Workspace:
    """Static buffers and prefilled argument structures of matcher + criterion for one (layers, batch, queries,
    matcher configuration) on one stream.  Capacities (target boxes, columns per problem) grow by re-building."""

    def __init__(self, dev, NL, B, Q, flat, matcher, eos_coef, s_cap, cols_cap):
        self.NL, self.B, self.Q, self.P = NL, B, Q, flat.P
        self.ppv, self.rows, self.per_frame = flat.problems_per_video, flat.rows_per_problem, flat.per_frame
        self.s_cap, self.cols_cap = s_cap, cols_cap
        self.views = static_views(torch.empty(0, dtype=torch.uint8), flat.P, B)      # sizes only
        self.tgt_buf = torch.zeros(self.views["n_fixed"] + 16 * s_cap, device=dev, dtype=torch.uint8)
        self.views = static_views(self.tgt_buf, flat.P, B)
        self.pred_idx = torch.zeros((NL, s_cap), device=dev, dtype=torch.int64)      # K <= S: row pitch = capacity
        self.tgt_idx = torch.zeros((NL, s_cap), device=dev, dtype=torch.int64)
        self.status = torch.zeros(2, device=dev, dtype=torch.int32)                   # zeroed ONCE; then owned by the kernels
        self.losses = torch.zeros((NL, 4), device=dev, dtype=torch.float32)
        # criterion partial sums + arrival counters (zeroed once; the kernels reset the counters themselves)
        self.scratch = torch.zeros(int(_lib.get_lib().svol_criterion_scratch_bytes(NL, B)), device=dev, dtype=torch.uint8)
        self.cost_ws = None
        if self.rows * cols_cap * 4 > COST_SMEM_LIMIT:        # blocks too large for shared memory: solved from HBM
            self.cost_ws = torch.empty((NL, self.rows * s_cap), device=dev, dtype=torch.float32)
        self.weights = (matcher.cost_class, matcher.cost_bbox, matcher.cost_giou)
        self.eos_coef = float(eos_coef)
        self.ma, self.ca = _lib.MatchArgs(), _lib.CriterionArgs()
        self.bound = None           # (logits ptr, boxes ptr) the argument structures currently point at
        self.loaded = None          # the FlatTargets whose bytes are in tgt_buf
        self.graph = None
        self.graph_key = None

    def fits(self, flat) -> bool:
        return flat.S <= self.s_cap and flat.max_cols <= self.cols_cap and flat.P == self.P

    def load(self, flat) -> None:
        if self.loaded is not flat:
            self.tgt_buf[:flat.n_static].copy_(flat.packed[:flat.n_static], non_blocking=True)
            self.loaded = flat

    def bind(self, logits, boxes) -> None:
        key = (logits.data_ptr(), boxes.data_ptr())
        if key == self.bound:
            return
        v = self.views
        fill_match_args(self.ma, logits, boxes, v, self.NL, self.B, self.Q, self.s_cap, self.ppv, self.rows, self.cols_cap,
                        self.per_frame, *self.weights, self.cost_ws, self.pred_idx, self.tgt_idx, self.status)
        a = self.ca
        a.logits, a.boxes, a.tgt_boxes = logits.data_ptr(), boxes.data_ptr(), v["tgt_boxes"].data_ptr()
        a.pred_idx, a.tgt_idx = self.pred_idx.data_ptr(), self.tgt_idx.data_ptr()
        a.match_video, a.video_tgt_off = None, v["video_tgt_off"].data_ptr()
        a.video_match_off, a.meta = v["video_match_off"].data_ptr(), v["meta"].data_ptr()
        a.losses = self.losses.data_ptr()
        a.NL, a.B, a.Q, a.K, a.idx_pitch = self.NL, self.B, self.Q, 0, self.s_cap
        a.eos_coef = self.eos_coef
        a.scratch = self.scratch.data_ptr()
        self.bound = key

    def launch(self, stream: int) -> None:
        lib = _lib.get_lib()
        _lib.check(lib.svol_match(C.byref(self.ma), stream), "match")
        _lib.check(lib.svol_criterion(C.byref(self.ca), stream), "criterion")

    def run(self, logits, boxes, use_graph: bool) -> None:
        """Matcher + criterion on the loaded targets.  ``use_graph``: the predictions live at stable addresses (the
        head's static output buffers), so the three launches are captured once and replayed."""
        self.bind(logits, boxes)
        if not use_graph:
            self.launch(_lib.stream_ptr())
            return
        if self.graph is None or self.graph_key != self.bound:
            self.launch(_lib.stream_ptr())              # warm-up outside capture (function attributes, module load)
            torch.cuda.current_stream().synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.launch(_lib.stream_ptr())
            self.graph, self.graph_key = g, self.bound
        self.graph.replay()

    def backward(self, grad_w, g_logits, g_boxes) -> None:
        _lib.check(_lib.get_lib().svol_criterion_backward(C.byref(self.ca), grad_w.data_ptr(), g_logits.data_ptr(),
                                                          g_boxes.data_ptr(), _lib.stream_ptr()), "criterion_backward")


class _CriterionFn(torch.autograd.Function):
    """losses [NL,4] = (loss_label, class_error, loss_bbox, loss_giou); differentiable in logits / boxes."""

    @staticmethod
    def forward(ctx, logits, boxes, ws, use_graph):
        ws.run(logits, boxes, use_graph)
        ctx.save_for_backward(logits, boxes)
        ctx.ws, ctx.loaded = ws, ws.loaded
        return ws.losses.clone()

    @staticmethod
    def backward(ctx, grad):
        logits, boxes = ctx.saved_tensors
        ws = ctx.ws
        if ws.loaded is not ctx.loaded or ws.bound != (logits.data_ptr(), boxes.data_ptr()):
            raise RuntimeError("criterion backward after the workspace was reused by another batch: call backward() "
                               "before the next criterion(...) of the same shape on this stream (train.py:222-229 does)")
        grad_w = grad[:, [0, 2, 3]].contiguous().float()           # class_error carries no gradient
        g_logits, g_boxes = torch.empty_like(logits), torch.empty_like(boxes)
        ws.backward(grad_w, g_logits, g_boxes)
        return g_logits, g_boxes, None, None


def _stack_layers(outputs):
    """([NL,B,Q,2], [NL,B,Q,4], static) in decoder-layer order (aux 0..NL-2, then the final layer).  Zero-copy when the
    dict came from ``svol_b200``'s SVANet (all entries are views of one buffer); ``static`` then tells whether those
    buffers are the inference engine's plan-owned outputs (stable addresses: the launches can be graph-replayed)."""
    lg, bx = outputs["pred_logits"], outputs["pred_boxes"]
    aux = outputs.get("aux_outputs", [])
    n = len(aux) + 1
    base_l, base_b = getattr(lg, "_base", None), getattr(bx, "_base", None)
    if (base_l is not None and base_b is not None and base_l.dim() == 4 and base_l.shape[0] == n
            and base_l.is_contiguous() and base_b.is_contiguous() and base_l.dtype == torch.float32
            and lg.data_ptr() == base_l[n - 1].data_ptr() and bx.data_ptr() == base_b[n - 1].data_ptr()
            and all(a["pred_logits"].data_ptr() == base_l[i].data_ptr() and a["pred_boxes"].data_ptr() == base_b[i].data_ptr()
                    for i, a in enumerate(aux))):
        return base_l, base_b, bool(getattr(base_l, "_svol_static", False))
    logits = torch.stack([a["pred_logits"] for a in aux] + [lg]).float().contiguous()
    boxes = torch.stack([a["pred_boxes"] for a in aux] + [bx]).float().contiguous()
    return logits, boxes, False


class SetCriterion(nn.Module):
    """Same constructor, attributes (``weight_dict``, ``empty_weight``) and output keys as loss.py:10-157.

    Without autograd (``torch.no_grad()`` / predictions that do not require grad) the returned 0-d tensors are views of
    a static buffer that the next call of the same shape on the same stream overwrites -- like the head's own outputs;
    read them (``float(v)``, a D2H copy) before that call, as train.py / test.py do."""

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
        self.use_graph = True
        self._ws = {}                     # (device, stream, NL, B, Q) -> _Workspace
        self._last = None                 # (workspace, flat) of the most recent call
        want = {"labels": (0, 1), "boxes": (2, 3)}
        self._cols = [c for name in self.losses for c in want[name]]
        self._keys = {}                   # NL -> [(key, flat index into losses.view(-1))]

    def _workspace(self, logits, flat) -> _Workspace:
        NL, B, Q = logits.shape[:3]
        key = (logits.device, _lib.stream_ptr(), NL, B, Q)
        ws = self._ws.get(key)
        if ws is None or not ws.fits(flat) or ws.eos_coef != float(self.eos_coef) or \
                ws.weights != (self.matcher.cost_class, self.matcher.cost_bbox, self.matcher.cost_giou):
            s_cap = _round_up(max(2 * flat.S, 256), 256)
            cols_cap = max(_round_up(flat.max_cols, 8), 8 if flat.per_frame else 64)
            if ws is not None:
                s_cap, cols_cap = max(s_cap, ws.s_cap), max(cols_cap, ws.cols_cap)
            ws = _Workspace(logits.device, NL, B, Q, flat, self.matcher, self.eos_coef, s_cap, cols_cap)
            self._ws[key] = ws
        return ws

    def forward(self, outputs, targets):
        logits, boxes, static = _stack_layers(outputs)
        if not logits.is_cuda:
            raise RuntimeError("svol_b200 criterion needs CUDA tensors; there is no CPU fallback")
        _lib.require_device()
        flat = self.matcher._flat(targets, logits.device, logits.shape[2])
        ws = self._workspace(logits, flat)
        ws.load(flat)
        self._last = (ws, flat)
        n = logits.shape[0]
        if torch.is_grad_enabled() and (logits.requires_grad or boxes.requires_grad):
            L = _CriterionFn.apply(logits, boxes, ws, False)
        else:
            ws.run(logits, boxes, static and self.use_graph)
            L = ws.losses
        keys = self._keys.get(n)
        if keys is None:
            keys = []
            for li in [n - 1] + list(range(n - 1)):
                suffix = "" if li == n - 1 else f"_{li}"
                keys += [(self.LOSS_NAMES[c] + suffix, li * 4 + c) for c in self._cols]
            self._keys[n] = keys
        flat_l = L.view(-1).unbind(0)
        return {k: flat_l[i] for k, i in keys}

    @property
    def last_indices(self):
        """(pred_idx [NL,K], tgt_idx [NL,K], FlatTargets) of the most recent call, on the device."""
        if self._last is None:
            return None
        ws, flat = self._last
        return ws.pred_idx[:, :flat.K], ws.tgt_idx[:, :flat.K], flat

    @property
    def last_status(self):
        return None if self._last is None else self._last[0].status

    def check_status(self) -> None:
        """Raises like scipy would have (NaN / -inf cost entries).  Synchronises the device."""
        if self._last is not None:
            _check_status(self._last[0].status)

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
