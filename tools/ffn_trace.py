#!/usr/bin/env python
"""Per-chunk timeline of the fused FFN kernel (debug build only):
    touch svol_b200/csrc/ffn_tc.cu; SVOL_EXTRA_NVCC_FLAGS=-DSVOL_FFN_TRACE bash svol_b200/csrc/build.sh
    python tools/ffn_trace.py"""
import ctypes as C
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from svol_b200 import _lib, ops

dev = torch.device("cuda:0")
B, L, d, ff = 32, 1568, 256, 2048
g = torch.Generator(device="cpu").manual_seed(0)
rnd = lambda *s: torch.randn(*s, generator=g)
M = int(os.environ.get("SVOL_FFN_TRACE_M", B * L))      # 10240 = the object-query FFN (80 tiles, one wave, query_pos table)
x = rnd(M, d).to(torch.bfloat16).to(dev)
w1 = (rnd(ff, d) / math.sqrt(d)).to(torch.bfloat16).to(dev)
w2 = (rnd(d, ff) / math.sqrt(ff)).to(torch.bfloat16).to(dev)
b1, b2 = rnd(ff).to(dev), rnd(d).to(dev)
ln = (torch.ones(d, device=dev), torch.zeros(d, device=dev))
if M == B * L:
    theta = ops.posenc_theta(torch.ones(B, L, device=dev)).reshape(-1)
    run = lambda: ops.ffn(x, w1, b1, w2, b2, ln, pos_theta=theta)
else:
    qpos = rnd(M // B, d).to(dev)
    run = lambda: ops.ffn(x, w1, b1, w2, b2, ln, pos=qpos, pos_mod=M // B)
for _ in range(2):
    run()
torch.cuda.synchronize()
buf = np.zeros((3, 64, 8), dtype=np.int64)
assert _lib.get_lib().svol_debug_ffn_trace(C.c_void_p(buf.ctypes.data)) == 0
t0 = buf[buf > 0].min()
print("epilogue warp 4, per chunk:   top  hacc_full   acc->reg  gelu done     h_free     stored")
for i in range(24):
    if buf[0, i, 0] > 0 or i == 0:
        print(f" chunk {i:2d} " + " ".join(f"{(v - t0) if v > 0 else -1:10d}" for v in buf[0, i, :6]))
print("MMA issuer, per chunk:   mma1 start mma1 issued    h_ready mma2 issued")
for i in range(24):
    print(f" chunk {i:2d} " + " ".join(f"{(v - t0) if v > 0 else -1:10d}" for v in buf[1, i, :4]))
print("tile epilogue (first epilogue warp), per tile:  oacc_full  O->reg  +res,x_free  LN done  store1 issued  sincos done(+read wait)  store2 issued  read wait")
for i in range(3):
    print(f" tile {i:2d} " + " ".join(f"{(v - t0) if v > 0 else -1:10d}" for v in buf[2, i, :8]))
