#!/usr/bin/env python
"""Kernel-level phases of attention CTA 0 (debug build -DSVOL_ATTN_TRACE): kernel entry -> set-up barrier -> register re-split ->
first score tile, and last probabilities stored -> last P V landed -> merged / stored -> every role done, in clk.
    python tools/attn_ends_trace.py [attn_self|attn_cross]"""
import ctypes as C, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from svol_b200 import _lib, ops
which = sys.argv[1] if len(sys.argv) > 1 else "attn_self"
dev = torch.device("cuda:0")
B, L, Q, H, d = 32, 1568, 320, 8, 256
Lq, Lk = (L, L) if which == "attn_self" else (Q, L)
g = torch.Generator(device="cpu").manual_seed(0)
q = (torch.randn(B * Lq, d, generator=g) * math.log2(math.e) / math.sqrt(32)).to(torch.bfloat16).to(dev)
k = torch.randn(B * Lk, d, generator=g).to(torch.bfloat16).to(dev)
vt = torch.randn(B * d, (Lk + 7) // 8 * 8, generator=g).to(torch.bfloat16).to(dev)
for _ in range(3):
    ops.attention(q, k, vt, B, H, Lq, Lk)
torch.cuda.synchronize()
buf = np.zeros((6, 64, 8), dtype=np.int64)
assert _lib.get_lib().svol_debug_attn_trace(C.c_void_p(buf.ctypes.data)) == 0
t0 = buf[0, 60, 3]
n = (Lk + 127) // 128
print(f"{which}: CTA 0, clk after kernel entry (warp 0)")
print(f"  set-up started {buf[0,60,0]-t0}, set-up barrier passed {buf[0,60,1]-t0}, registers re-split {buf[0,60,2]-t0}")
for gidx in range(4):
    rows = [j for j in range(n) if buf[gidx, j, 7] > 0]
    print(f"  softmax g{gidx}: first top {buf[gidx,0,0]-t0}, first S {buf[gidx,0,1]-t0}, last P stored {buf[gidx,rows[-1],7]-t0} (tile {rows[-1]}), "
          f"O in registers {buf[gidx,61,0]-t0}, merged / stored {buf[gidx,61,1]-t0}, all roles done {buf[gidx,61,2]-t0}")
