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

    def forward(self, src_sketch, src_video):
        N, T = src_video.shape[:2]
        f = self.video_trunk(src_video.flatten(0, 1))                   # (N*T, C, h, w)
        f = f.flatten(2).transpose(1, 2).reshape(N, -1, f.shape[1])     # (N, T*h*w, C)
        s = self.sketch_trunk(src_sketch.flatten(0, 1)).flatten(1).reshape(N, -1, f.shape[-1])
        return s, f


class SketchLocalizationModel(nn.Module):
    def __init__(self, backbone, head):
        super().__init__()
        self.backbone = backbone
        self.head = head

    def forward(self, src_sketch, src_video, src_sketch_mask=None, src_video_mask=None):
        N, T = src_video.shape[:2]
        src_sketch, src_video = self.backbone(src_sketch, src_video)
        # model.py:21-22: per-frame masks repeated over each frame's tokens
        src_sketch_mask = src_sketch_mask.repeat_interleave(src_sketch.shape[1], dim=1)
        src_video_mask = src_video_mask.repeat_interleave(src_video.shape[1] // T, dim=1)
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
