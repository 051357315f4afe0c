#!/usr/bin/env python
"""Input LayerNorm (fp32 frame features -> bf16) at the C2 size, operands flushed out of L2 before every launch:
time per launch and achieved HBM bandwidth (103 MB read + 51 MB written).   python tools/ln_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from svol_b200 import ops

dev = torch.device("cuda:0")
rows, cols = 32 * 1568, 512
x = torch.randn(rows, cols, device=dev)
w, b = torch.ones(cols, device=dev), torch.zeros(cols, device=dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
ts = []
for rep in range(25):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    y = ops.layernorm_to_bf16(x, w, b)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
ts = sorted(ts[5:])
med = ts[len(ts) // 2]
ref = torch.nn.functional.layer_norm(x, (cols,), w, b)
err = (y.float() - ref).abs().max().item()
print(f"ln_in: {med:.1f} us per launch (median of 20), {(rows * cols * 6) / med / 1e3:.0f} GB/s, max |err| vs torch {err:.4f}")
