"""SVANet head, drop-in for lib/modeling/svanet.py, computed by the sm_100a launch plan."""
from __future__ import annotations

import torch
from torch import nn

from ..engine import HeadEngine
from ..train_engine import TrainEngine
from .cross_modal_transformer import build_cross_modal_transformer
from .position_encoding import build_position_encoding


class MLP(nn.Module):
    """Parameters of the box head (svanet.py:144-156): Linear layers under ``layers.{i}``."""

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        h = [hidden_dim] * (num_layers - 1)
        self.layers = nn.ModuleList(nn.Linear(n, k) for n, k in zip([input_dim] + h, h + [output_dim]))


class LinearLayer(nn.Module):
    """Parameters of one input-projection stage (svanet.py:159-181): ``LayerNorm`` and ``net.1``."""

    def __init__(self, in_hsz, out_hsz, layer_norm=True, dropout=0.1, relu=True):
        super().__init__()
        if not layer_norm:
            raise NotImplementedError("the fused input projection always applies LayerNorm (the reference does too)")
        self.relu = relu
        self.layer_norm = layer_norm
        self.LayerNorm = nn.LayerNorm(in_hsz)
        self.net = nn.Sequential(nn.Dropout(dropout), nn.Linear(in_hsz, out_hsz))


class _HeadTrainFn(torch.autograd.Function):
    """Training-mode forward / backward of the head as ONE autograd node, so that the reference's
    ``loss.backward()`` (train.py:229) drives the CUDA backward plan.  Inputs after the four data tensors are the
    module's parameters (their gradients are the node's outputs).  Features that require a gradient -- a backbone trained
    through the head, as train.py:72 does -- get d(loss)/d(src_video), d(loss)/d(src_sketch) as well."""

    @staticmethod
    def forward(ctx, module, src_sketch, src_sketch_mask, src_video, src_video_mask, *params):
        eng = module.train_engine
        ctx.want_sketch, ctx.want_video = bool(ctx.needs_input_grad[1]), bool(ctx.needs_input_grad[3])
        ctx.sketch_shape = tuple(src_sketch.shape)
        logits, boxes = eng.forward(src_sketch, src_sketch_mask, src_video, src_video_mask, want_input_grads=ctx.want_video)
        ctx.module, ctx.n_params, ctx.token = module, len(params), eng.forward_token
        return logits.clone(), boxes.clone()

    @staticmethod
    def backward(ctx, g_logits, g_boxes):
        module = ctx.module
        eng = module.train_engine
        eng.backward(g_logits.contiguous().float(), g_boxes.contiguous().float(), token=ctx.token)
        g_sketch = g_video = None
        if ctx.want_video or ctx.want_sketch:
            g_sketch, g_video = eng.input_grads()
            g_sketch = g_sketch.view(ctx.sketch_shape) if ctx.want_sketch else None
            g_video = g_video if ctx.want_video else None
        if not eng.publish_grads:        # fused-optimizer loop: gradients stay in eng.grad_flat (FusedAdamW.step(from_engine=True))
            return (None, g_sketch, None, g_video, None) + (None,) * ctx.n_params
        unused = {id(p) for p in module.params_without_grad()}            # None in the reference's autograd too
        grads = [eng.grad_of(p) if (p.requires_grad and id(p) not in unused) else None for p in module.parameters()]
        return (None, g_sketch, None, g_video, None, *grads)


class SVANet(nn.Module):
    """Same constructor, parameters and outputs as the reference's SVANet (svanet.py:14-141)."""

    def __init__(self, transformer, sketch_position_embed, video_position_embed, input_vid_dim, input_skch_dim,
                 num_queries, input_dropout=0.1, aux_loss=True, use_sketch_pos=True, n_input_proj=2, num_classes=2,
                 vis_mode=None, use_graph=True):
        super().__init__()
        self.num_queries = num_queries
        self.num_classes = num_classes
        self.transformer = transformer
        self.sketch_position_embed = sketch_position_embed
        self.video_position_embed = video_position_embed
        hidden_dim = transformer.d_model
        self.bbox_embed = MLP(hidden_dim, hidden_dim, 4, 3)
        self.use_sketch_pos = use_sketch_pos
        self.class_embed = nn.Linear(hidden_dim, 2)
        self.n_input_proj = n_input_proj
        self.class_head = nn.Linear(hidden_dim, num_classes)       # unused by forward, kept for checkpoints
        self.query_embed = nn.Embedding(num_queries, hidden_dim)
        relu_args = [True] * 3
        relu_args[n_input_proj - 1] = False
        dims_v = [input_vid_dim, hidden_dim, hidden_dim]
        dims_s = [input_skch_dim, hidden_dim, hidden_dim]
        self.input_video_proj = nn.Sequential(*[
            LinearLayer(dims_v[i], hidden_dim, layer_norm=True, dropout=input_dropout, relu=relu_args[i])
            for i in range(n_input_proj)])
        self.input_sketch_proj = nn.Sequential(*[
            LinearLayer(dims_s[i], hidden_dim, layer_norm=True, dropout=input_dropout, relu=relu_args[i])
            for i in range(n_input_proj)])
        self.vis_mode = vis_mode
        self.aux_loss = aux_loss
        self.input_dropout = input_dropout
        self._engine = HeadEngine(self, use_graph=use_graph)
        self._train_engine = None

    def params_without_grad(self):
        """Parameters the forward never reaches -- ``.grad`` stays None in the reference's autograd as well:
        ``class_head`` (svanet.py:125 uses class_embed) and the output projection of the sketch->video attention, whose
        attended values are discarded (cross_modal_transformer.py:124: only the attention weights are used)."""
        out = list(self.class_head.parameters())
        for layer in self.transformer.layers:
            out += list(layer.sketch_video_cross_attn.out_proj.parameters())
        return out

    @property
    def engine(self) -> HeadEngine:
        return self._engine

    @property
    def train_engine(self) -> TrainEngine:
        if self._train_engine is None:
            self._train_engine = TrainEngine(self)
        return self._train_engine

    def forward(self, src_sketch, src_sketch_mask, src_video, src_video_mask):
        """src_sketch (B,1,D_s), src_sketch_mask (B,1), src_video (B,L,D_v), src_video_mask (B,L) float {0,1}.
        Inference also accepts src_video as the backbone's (B,T,D_v,h,w) feature map (L = T*h*w tokens in (frame,
        position) order): the first LayerNorm then reads the channel-major layout directly.
        Returns {'pred_logits' (B,Q,2), 'pred_boxes' (B,Q,4) cxcywh, 'aux_outputs': [...]} (svanet.py:128-141)."""
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            # train.py:216-232: the forward keeps what the backward needs; loss.backward() runs the CUDA backward plan
            if src_video.dim() == 5:        # training plan takes token rows: (B,T,C,h,w) -> (B, T*h*w, C)
                src_video = src_video.flatten(3).permute(0, 1, 3, 2).reshape(src_video.shape[0], -1, src_video.shape[2])
            logits, boxes = _HeadTrainFn.apply(self, src_sketch, src_sketch_mask, src_video, src_video_mask,
                                               *self.parameters())
        else:
            logits, boxes = self._engine.forward(src_sketch, src_sketch_mask, src_video, src_video_mask)
        out = {"pred_logits": logits[-1], "pred_boxes": boxes[-1]}
        if self.aux_loss:
            out["aux_outputs"] = [{"pred_logits": a, "pred_boxes": b} for a, b in zip(logits[:-1], boxes[:-1])]
        if self.vis_mode is not None:
            if logits.requires_grad:
                raise NotImplementedError("vis_mode returns detached decoder states; use it under torch.no_grad()")
            key = tuple(src_video.shape) if src_video.dim() == 3 else \
                (src_video.shape[0], src_video.shape[1] * src_video.shape[3] * src_video.shape[4], src_video.shape[2])
            hs = self._engine.plan_for(*key).buf["hs"]
            return out, hs.float().view(hs.shape[0], src_video.shape[0], self.num_queries, -1)
        return out


def build_svanet(args):
    """svanet.py:184-200."""
    transformer = build_cross_modal_transformer(args)
    sketch_position_embed, video_position_embed = build_position_encoding(args)
    return SVANet(
        transformer, sketch_position_embed, video_position_embed,
        input_vid_dim=args.input_vid_dim, input_skch_dim=args.input_skch_dim, num_queries=args.num_queries,
        input_dropout=args.input_dropout, aux_loss=args.aux_loss, use_sketch_pos=args.use_sketch_pos,
        n_input_proj=args.n_input_proj, vis_mode=args.vis_mode,
        use_graph=bool(getattr(args, "use_cuda_graph", True)),
    )
