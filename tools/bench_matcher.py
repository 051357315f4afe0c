#!/usr/bin/env python
"""Matcher / criterion kernels alone at the headline (C2) and stress (C5) sizes: device time per call, algorithmic
bytes (logits + boxes of all layers, targets, offsets, indices, block-diagonal costs) and the implied GB/s,
assignment problems per second.   python tools/bench_matcher.py"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dataclasses import replace
import numpy as np
import torch
from svol_b200 import synth
from svol_b200.modeling import build_loss

dev = torch.device("cuda:0")
peak = 6538.0
if os.path.exists("MEASURED_PEAKS.json"):
    peak = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", peak)
cases = [("C2 per-frame (10 x n_f<=2)", synth.CONFIGS["C2"], 32, 2),
         ("C5 per-frame (100 x n_f<=50)", synth.CONFIGS["C5"], 32, 50),
         ("C2 video matcher (320 x n_v)", replace(synth.CONFIGS["C2"], matcher="video_matcher"), 32, 2),
         ("C5-video (320 x n_v<=320)", replace(synth.CONFIGS["C2"], matcher="video_matcher"), 32, 10)]
for name, cfg, B, mpf in cases:
    crit = build_loss(cfg.to_namespace()).to(dev)
    lg, bx = synth.make_predictions(cfg, B, 0)
    out = {"pred_logits": torch.from_numpy(lg[-1]).to(dev), "pred_boxes": torch.from_numpy(bx[-1]).to(dev),
           "aux_outputs": [{"pred_logits": torch.from_numpy(a).to(dev), "pred_boxes": torch.from_numpy(b).to(dev)}
                           for a, b in zip(lg[:-1], bx[:-1])]}
    tg = synth.targets_to_torch(synth.make_targets(cfg, B, 0, max_per_frame=mpf))
    with torch.no_grad():
        for _ in range(3):
            crit(out, tg)
        torch.cuda.synchronize()
        # the three launches of one call are captured in a CUDA graph so that the device time is measured, not the
        # Python / ctypes issue time (~0.12 ms per call, which hides under the forward in the real step)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            crit(out, tg)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    flat = crit.last_indices[2]
    NL, Q = lg.shape[0], lg.shape[2]
    bytes_alg = NL * B * Q * 24 + flat.S * 16 + (flat.P + 1) * 4 + 2 * NL * flat.K * 8 + 4 * NL * 4 + 2 * 4 * NL * flat.cost_total
    print(f"{name:32s} problems/call {NL * flat.P:6d}  boxes {flat.S:6d}  matched/layer {flat.K:6d}  "
          f"{ms * 1e3:8.1f} us/call  {NL * flat.P / ms / 1e3:8.2f} M problems/s  algorithmic {bytes_alg / 1e6:7.2f} MB -> "
          f"{bytes_alg / ms / 1e6:7.1f} GB/s ({100 * bytes_alg / ms / 1e6 / peak:4.1f} % of {peak:.0f})")
