#!/usr/bin/env python
"""Where a step's time goes: host time per call vs device time, forward only vs forward + criterion."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dataclasses import replace
from svol_b200 import synth
from svol_b200.modeling import build_loss, build_svanet

dev = torch.device("cuda:0")
cfg = replace(synth.CONFIGS["C2"], num_layers=2)
B = 32
ns = cfg.to_namespace()
model = build_svanet(ns)
model.load_state_dict({k: torch.from_numpy(v) for k, v in synth.random_state_dict(cfg, 0).items()}, strict=True)
model = model.to(dev).eval()
criterion = build_loss(ns).to(dev)
inp = synth.make_inputs(cfg, B, seed=0, padded=True)
tg = synth.targets_to_torch(synth.make_targets(cfg, B, seed=0, frame_mask=inp["frame_mask"]))
d = {k: torch.from_numpy(inp[k]).to(dev) for k in ("src_sketch", "src_sketch_mask", "src_video", "src_video_mask")}

def fwd():
    return model(d["src_sketch"], d["src_sketch_mask"], d["src_video"], d["src_video_mask"])

def full():
    return criterion(fwd(), tg)

def measure(fn, name, n=50):
    with torch.no_grad():
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        t_host = time.perf_counter() - t0
        torch.cuda.synchronize()
    print(f"{name:28s} device {e0.elapsed_time(e1) / n:7.3f} ms/step   host issue {t_host / n * 1e3:7.3f} ms/step")

measure(fwd, "forward (graph)")
measure(full, "forward + criterion")
plan = model.engine.plan_for(B, cfg.video_len, cfg.input_vid_dim)
measure(lambda: plan.graph.replay(), "graph replay only")
out = fwd()
measure(lambda: criterion(out, tg), "criterion only")
model.engine.use_graph = False
measure(fwd, "forward (eager launches)")
