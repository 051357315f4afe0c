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
import ctypes as C
from svol_b200 import _lib
from svol_b200.modeling.matcher import fill_match_args


def time_graph(fn, n=20):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for name, cfg, B, mpf in cases:
    crit = build_loss(cfg.to_namespace()).to(dev)
    lg, bx = synth.make_predictions(cfg, B, 0)
    logits, boxes = torch.from_numpy(lg).to(dev), torch.from_numpy(bx).to(dev)
    out = {"pred_logits": logits[-1], "pred_boxes": boxes[-1],
           "aux_outputs": [{"pred_logits": a, "pred_boxes": b} for a, b in zip(logits[:-1], boxes[:-1])]}
    tg = synth.targets_to_torch(synth.make_targets(cfg, B, 0, max_per_frame=mpf))
    with torch.no_grad():
        for _ in range(3):
            crit(out, tg)
        torch.cuda.synchronize()
        crit.check_status()
        # the launches of one call are captured in a CUDA graph so that the device time is measured, not the
        # Python / ctypes issue time
        ms = time_graph(lambda: crit(out, tg))
        flat = crit.last_indices[2]
        NL, Q = lg.shape[0], lg.shape[2]
        m = crit.matcher
        ws = torch.empty((NL, max(flat.cost_total, 1)), device=dev, dtype=torch.float32)
        pi = torch.empty((NL, flat.K), device=dev, dtype=torch.int64); ti = torch.empty_like(pi)
        st = torch.zeros(2, device=dev, dtype=torch.int32)
        part = {}
        for label, mode, solver, cws in (("fused", 0, 0, None), ("fused, smem-state solver", 0, 1, None), ("cost blocks only", 1, 0, ws),
                                         ("solve only", 2, 0, ws), ("solve only, smem-state solver", 2, 1, ws)):
            if cws is None and flat.rows_per_problem * flat.max_cols * 4 > 96 * 1024:
                cws = ws
            a = fill_match_args(_lib.MatchArgs(), logits, boxes, flat, NL, B, Q, flat.K, flat.problems_per_video,
                                flat.rows_per_problem, flat.max_cols, flat.per_frame, m.cost_class, m.cost_bbox, m.cost_giou,
                                cws, pi, ti, st, mode, solver)
            part[label] = time_graph(lambda: _lib.check(_lib.get_lib().svol_match(C.byref(a), _lib.stream_ptr()), "match")) * 1e3
    bytes_in = NL * B * Q * 24 + flat.S * 16 + (flat.P + 1) * 4 + 2 * NL * flat.K * 8 + 4 * NL * 4
    bytes_cost = 4 * NL * flat.cost_total
    print(f"{name:32s} problems/call {NL * flat.P:6d}  boxes {flat.S:6d}  matched/layer {flat.K:6d}  max cols {flat.max_cols:4d}\n"
          f"    criterion call (match + finalize + criterion) {ms * 1e3:8.1f} us  = {NL * flat.P / ms / 1e3:8.2f} M problems/s\n"
          + "".join(f"    svol_match {k:34s} {v:8.1f} us\n" for k, v in part.items())
          + f"    algorithmic bytes: inputs + indices {bytes_in / 1e6:6.2f} MB; cost blocks {bytes_cost / 1e6:6.2f} MB written -> cost-only pass "
            f"{bytes_cost / part['cost blocks only'] / 1e3:7.1f} GB/s ({100 * bytes_cost / part['cost blocks only'] / 1e3 / peak:4.1f} % of {peak:.0f})")
