"""Timings of the SURVEY 8(f) rows at C2 shapes (B=32): backbone hand-off, evaluation metrics, target packing.
usage: python tools/bench_next_rows.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from svol_b200 import _lib, ops, synth
from svol_b200.evaluate import SVOLEvaluator, flatten_eval_targets
from svol_b200.modeling import postprocess, targets as T
from oracle import eval_oracle as ev

dev = "cuda:0"
cfg = synth.CONFIGS["C2"]
B, Tn, C, hw = 32, cfg.num_frames, cfg.input_vid_dim, 7


def gpu_time(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


# ---- 8f-2 backbone hand-off
fmap = torch.randn(B, Tn, C, hw, hw, device=dev).relu_()
w, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
y = torch.empty(B * Tn * hw * hw, C, device=dev, dtype=torch.bfloat16)
lib, P, st = _lib.get_lib(), _lib.ptr, _lib.stream_ptr
fused = lambda: _lib.check(lib.svol_layernorm_nchw_to_bf16(P(fmap), P(w), P(b), P(y), B * Tn, C, hw * hw, 1e-5, st()), "nchw")
def unfused():
    tok = fmap.flatten(3).permute(0, 1, 3, 2).reshape(B, -1, C)           # the reference's reshape / transpose copy
    return ops.layernorm_to_bf16(tok.reshape(-1, C), w, b)
mb = fmap.numel() * 4 / 1e6
t_f, t_u = gpu_time(fused), gpu_time(unfused)
print(f"8f-2 hand-off  ({mb:.0f} MB feature map): fused LayerNorm-from-NCHW {t_f:.1f} us ({(mb + mb / 2) / t_f * 1e3 / 1e3:.2f} TB/s of "
      f"{mb * 1.5:.0f} MB algorithmic), permute copy + LayerNorm {t_u:.1f} us")

# ---- 8f-3 evaluation metrics
logits, boxes = synth.make_eval_predictions(cfg, B, 0)
inp = synth.make_inputs(cfg, B, 0, padded=True)
tg = synth.make_targets(cfg, B, 0, frame_mask=inp["frame_mask"])
lg, bx = torch.from_numpy(logits).to(dev), torch.from_numpy(boxes).to(dev)
post, _ = postprocess(lg, bx, cfg.num_frames)
evaluator = SVOLEvaluator(cfg.num_frames, cfg.num_queries_per_frame)
def eval_gpu():
    evaluator.max1.clear(); evaluator.max5.clear(); evaluator.ap.clear()
    evaluator.update(post, tg)
t_e = gpu_time(eval_gpu, reps=10)
t0 = time.perf_counter(); eval_gpu(); torch.cuda.synchronize(); m = evaluator.summary(); wall = time.perf_counter() - t0
gt, gt_off, frame_off, fidx = flatten_eval_targets(tg, cfg.num_frames)
pred = post.cpu().numpy().reshape(-1, cfg.num_queries_per_frame, 5)[fidx]
t0 = time.perf_counter(); ref = ev.eval_svol(pred, gt, gt_off, frame_off); cpu = time.perf_counter() - t0
print(f"8f-3 evaluation (B={B}: {len(fidx)} frames, {gt.shape[0]} boxes): GPU kernels + host marshalling {t_e:.0f} us per batch "
      f"({wall * 1e3:.2f} ms wall incl. summary), numpy oracle {cpu * 1e3:.0f} ms; identical: {m == ref}")

# ---- 8f-4 target packing
tt = synth.targets_to_torch(tg)
t0 = time.perf_counter()
for _ in range(20):
    packed = T.pack_targets(tt, True, cfg.num_frames, cfg.num_queries, cfg.num_queries_per_frame)
pack = (time.perf_counter() - t0) / 20
d = torch.device(dev)
def up_packed():
    T.flatten_targets(packed, d, True, cfg.num_frames, cfg.num_queries, cfg.num_queries_per_frame)
def up_walk():
    T.flatten_targets(tt, d, True, cfg.num_frames, cfg.num_queries, cfg.num_queries_per_frame)
for fn, name in ((up_packed, "packed at collate time"), (up_walk, "nested dicts walked in the training process")):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    print(f"8f-4 targets -> device, {name}: {(time.perf_counter() - t0) / 50 * 1e6:.0f} us per batch (host)")
print(f"8f-4 pack_targets in the DataLoader worker: {pack * 1e6:.0f} us per batch")
