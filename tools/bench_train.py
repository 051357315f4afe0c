"""Training-step timing of the CUDA head (config 3 of BASELINE.json on one GPU: forward in train mode +
PerFrameMatcher + SetCriterion + backward + fused AdamW), with a per-call-family breakdown of the two plans.
usage: python tools/bench_train.py [--batch 32] [--steps 10] [--breakdown]"""
import argparse, os, sys, time, re, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dataclasses import replace
import torch
from svol_b200 import synth, _lib
from svol_b200.modeling import build_svanet, build_loss
from svol_b200.optim import FusedAdamW

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--config", default="C2")
ap.add_argument("--breakdown", action="store_true")
args = ap.parse_args()
dev = "cuda:0"
cfg = synth.CONFIGS[args.config]        # input_dropout 0.4, the reference default
model = build_svanet(cfg.to_namespace())
sd = synth.random_state_dict(cfg, 0)
model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
model = model.to(dev).train()
criterion = build_loss(cfg.to_namespace()).to(dev).train()
opt = FusedAdamW(model, lr=1e-4, weight_decay=1e-4)
model.train_engine.publish_grads = False
B = args.batch
inp = synth.make_inputs(cfg, B, 0, padded=True)
targets = synth.targets_to_torch(synth.make_targets(cfg, B, 0, frame_mask=inp["frame_mask"]))
t = {k: torch.from_numpy(v).to(dev) for k, v in inp.items()}
wd = criterion.weight_dict

def step():
    out = model(t["src_sketch"], t["src_sketch_mask"], t["src_video"], t["src_video_mask"])
    loss_dict = criterion(out, targets)
    total = sum(loss_dict[k] * wd[k] for k in loss_dict if k in wd)
    total.backward()
    opt.step(from_engine=True)
    return total

for _ in range(3):
    l = step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(args.steps):
    l = step()
e1.record()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / args.steps * 1e3
ms = e0.elapsed_time(e1) / args.steps
print(f"train step {args.config} B={B}: {ms:.3f} ms device, {wall:.3f} ms wall, {B / ms * 1e3:.0f} pairs/s, loss {float(l):.4f}, "
      f"mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
if args.breakdown:
    plan = model.train_engine._last
    st = torch.cuda.current_stream().cuda_stream
    for which in ("fwd", "bwd"):
        calls = plan[which].calls
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(calls) + 1)]
        evs[0].record()
        for i, (name, fn, a) in enumerate(calls):
            _lib.check(fn(*a, st), name)
            evs[i + 1].record()
        torch.cuda.synchronize()
        fam = collections.OrderedDict()
        for i, (name, fn, a) in enumerate(calls):
            key = re.sub(r"^l\d+\.", "", name)
            fam.setdefault(key, [0, 0.0])
            fam[key][0] += 1
            fam[key][1] += evs[i].elapsed_time(evs[i + 1])
        tot = sum(v[1] for v in fam.values())
        print(f"--- {which}: {len(calls)} calls, {tot:.3f} ms")
        for k, (n, ms_) in sorted(fam.items(), key=lambda kv: -kv[1][1])[:45]:
            print(f"  {k:28s} x{n:<3d} {ms_:8.3f} ms  {100 * ms_ / tot:5.1f}%")
