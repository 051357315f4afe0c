#!/usr/bin/env python
"""Where the HOST time of one forward + criterion step goes (cProfile over the bench's resident-input step):
   python tools/host_profile.py [steps]"""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from svol_b200 import synth
from svol_b200.modeling import build_loss, build_svanet

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
dev = torch.device("cuda:0")
cfg = synth.CONFIGS["C2"]
B = 32
model = build_svanet(cfg.to_namespace())
model.load_state_dict({k: torch.from_numpy(v) for k, v in synth.random_state_dict(cfg, 0).items()})
model = model.to(dev).eval()
crit = build_loss(cfg.to_namespace()).to(dev)
inp = synth.make_inputs(cfg, B, 0, padded=True)
tg = synth.targets_to_torch(synth.make_targets(cfg, B, 0, frame_mask=inp["frame_mask"]))
bufs = model.engine.input_buffers(B, cfg.video_len, cfg.input_vid_dim)
bufs["src_video"].copy_(torch.from_numpy(inp["src_video"]))
bufs["src_sketch"].copy_(torch.from_numpy(inp["src_sketch"]).reshape(B, -1))
bufs["src_video_mask"].copy_(torch.from_numpy(inp["src_video_mask"]))
sk, sm = bufs["src_sketch"].view(B, 1, -1), torch.ones(B, 1, device=dev)


def step():
    out = model(sk, sm, bufs["src_video"], bufs["src_video_mask"])
    return crit(out, tg)


with torch.no_grad():
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"host enqueue per step (no profiler): {(t1 - t0) / steps * 1e6:.1f} us (device-bound if ~1350 us)")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(steps):
        step()
    pr.disable()
    torch.cuda.synchronize()
st = pstats.Stats(pr).sort_stats("cumulative")
st.print_stats(28)
