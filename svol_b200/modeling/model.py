"""SketchLocalizationModel wrapper, drop-in for lib/modeling/model.py:9-44.

The backbone (torchvision ResNet trunks, lib/modeling/backbone.py:65-89,133-152) is upstream of the
accelerated path and stays a cuDNN module; it is restated here only so ``build_model`` returns a
complete model.  The head is the CUDA ``SVANet``.
"""
from __future__ import annotations

import torch
from torch import nn

from .svanet import build_svanet


class ResNetBackbone(nn.Module):
    """Frames -> (N, T*h*w, C) tokens, sketch -> (N, 1, C) (backbone.py:65-89); random-init trunks unless the
    caller loads weights (the reference downloads ImageNet weights, which needs network access)."""

    def __init__(self, video_arch="resnet34", sketch_arch="resnet18"):
        super().__init__()
        import torchvision
        v = getattr(torchvision.models, video_arch)(weights=None)
        s = getattr(torchvision.models, sketch_arch)(weights=None)
        self.video_trunk = nn.Sequential(*list(v.children())[:-2])       # conv feature map (C, 7, 7)
        self.sketch_trunk = nn.Sequential(*list(s.children())[:-1])      # pooled (C, 1, 1)
        self.out_dim = v.fc.in_features

    def feature_map(self, src_sketch, src_video):
        """Sketch feature (N, 1, C) and the video trunk's output AS cuDNN LEAVES IT, (N, T, C, h, w): the CUDA head's
        first LayerNorm reads this layout directly (svol_layernorm_nchw_to_bf16), so the reference's reshape /
        transpose copy of the whole feature tensor (backbone.py:72-89) never happens.  (Reshaping with explicit N
        also avoids the reference's batch-1 ``.squeeze()`` collapse, backbone.py:78.)"""
        N, T = src_video.shape[:2]
        f = self.video_trunk(src_video.flatten(0, 1))                   # (N*T, C, h, w)
        s = self.sketch_trunk(src_sketch.flatten(0, 1)).flatten(1).reshape(N, -1, f.shape[1])
        return s, f.reshape(N, T, *f.shape[1:])

    def forward(self, src_sketch, src_video):
        s, f = self.feature_map(src_sketch, src_video)
        N = f.shape[0]
        f = f.flatten(3).permute(0, 1, 3, 2).reshape(N, -1, f.shape[2])  # (N, T*h*w, C)
        return s, f


class SketchLocalizationModel(nn.Module):
    def __init__(self, backbone, head):
        super().__init__()
        self.backbone = backbone
        self.head = head

    def forward(self, src_sketch, src_video, src_sketch_mask=None, src_video_mask=None):
        N, T = src_video.shape[:2]
        training = self.training and torch.is_grad_enabled()
        if hasattr(self.backbone, "feature_map") and not training:
            # fused hand-off: the head normalises the trunk's channel-major output in place
            src_sketch, src_video = self.backbone.feature_map(src_sketch, src_video)          # (N,1,C), (N,T,C,h,w)
            tokens_per_frame = src_video.shape[-1] * src_video.shape[-2]
        else:
            src_sketch, src_video = self.backbone(src_sketch, src_video)
            tokens_per_frame = src_video.shape[1] // T
        # model.py:21-22: per-frame masks repeated over each frame's tokens
        src_sketch_mask = src_sketch_mask.repeat_interleave(src_sketch.shape[1], dim=1)
        src_video_mask = src_video_mask.repeat_interleave(tokens_per_frame, dim=1)
        return self.head(src_sketch, src_sketch_mask, src_video, src_video_mask)


def build_model(args, backbone: nn.Module = None):
    """model.py:31-44.  ``backbone`` may be supplied by the caller (e.g. the reference's own
    ``build_backbone(args)``); otherwise a ResNet-34 / ResNet-18 pair is built."""
    if getattr(args, "sketch_head", "svanet") != "svanet":
        raise NotImplementedError
    if backbone is None:
        backbone = ResNetBackbone()
        args.input_vid_dim = args.input_skch_dim = backbone.out_dim      # backbone.py:140-141
    head = build_svanet(args)
    return SketchLocalizationModel(backbone, head)
