#!/usr/bin/env python
"""Runs the C2 (B=32) head forward + criterion eagerly (no CUDA graph) a few times; the LAST iteration sits in the NVTX range
"profiled" -- the target of `ncu --nvtx --nvtx-include "profiled/"` captures of every kernel of one step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from svol_b200 import synth
from svol_b200.modeling import build_loss, build_svanet

dev = torch.device("cuda:0")
cfg = synth.CONFIGS[os.environ.get("SVOL_CONFIG", "C2")]
B = int(os.environ.get("SVOL_BATCH", "32"))
ns = cfg.to_namespace()
ns.use_cuda_graph = False
model = build_svanet(ns)
model.load_state_dict({k: torch.from_numpy(v) for k, v in synth.random_state_dict(cfg, 0).items()})
model = model.to(dev).eval()
crit = build_loss(ns).to(dev)
crit.use_graph = False
inp = synth.make_inputs(cfg, B, 0, padded=True)
tg = synth.targets_to_torch(synth.make_targets(cfg, B, 0, frame_mask=inp["frame_mask"]))
t = {k: torch.from_numpy(inp[k]).to(dev) for k in ("src_sketch", "src_sketch_mask", "src_video", "src_video_mask")}
with torch.no_grad():
    for it in range(3):
        if it == 2:
            torch.cuda.synchronize()
            torch.cuda.nvtx.range_push("profiled")
        out = model(t["src_sketch"], t["src_sketch_mask"], t["src_video"], t["src_video_mask"])
        losses = crit(out, tg)
        if it == 2:
            torch.cuda.synchronize()
            torch.cuda.nvtx.range_pop()
print({k: round(float(v), 4) for k, v in losses.items()})
