"""Per-parameter gradient error of the CUDA head backward against the oracle's autograd (debug aid).
usage: python tools/debug_grads.py C1a 2 0"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dataclasses import replace
import torch
from svol_b200 import synth
from oracle import torch_port as tp
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_train_gpu as T

name, batch, seed = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
cfg = replace(synth.CONFIGS[name], input_dropout=0.0)
model, sd, inp, (gl, gb), grads, logits, boxes = T._head_grads(cfg, batch, seed)
ref, rl, rb = tp.head_gradients(tp.state_dict_to_torch(sd), inp["src_sketch"], inp["src_sketch_mask"], inp["src_video"],
                                inp["src_video_mask"], gl, gb, nheads=cfg.nheads)
print("fwd max err logits", float((logits - rl).abs().max()), "boxes", float((boxes - rb).abs().max()))
scale = max(float(v.double().norm()) for v in ref.values())
rows = []
for k, r in ref.items():
    g = grads[k].double(); r = r.double()
    rows.append((float((g - r).norm()) / max(float(r.norm()), 1e-4 * scale), float(r.norm()), float(g.norm()), k))
for e, rn, gn, k in sorted(rows, reverse=True)[:int(sys.argv[4]) if len(sys.argv) > 4 else 25]:
    print(f"{e:9.4f}  ref {rn:10.4g}  got {gn:10.4g}  {k}")
