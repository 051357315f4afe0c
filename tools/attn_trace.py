#!/usr/bin/env python
"""Phase timeline of the attention kernel (debug build only):
    SVOL_EXTRA_NVCC_FLAGS=-DSVOL_ATTN_TRACE bash svol_b200/csrc/build.sh   (after touching attn_tc.cu)
    python tools/attn_trace.py [attn_self|attn_cross]
CTA (0,0,0) records clock64() at the phase boundaries of every key tile; this prints them relative to the CTA's
first record, for the two softmax warpgroups (warp 0 / warp 4) and the MMA issuer."""
import ctypes as C
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from svol_b200 import _lib, ops

which = sys.argv[1] if len(sys.argv) > 1 else "attn_self"
dev = torch.device("cuda:0")
B, L, Q, H, d = 32, 1568, 320, 8, 256
g = torch.Generator(device="cpu").manual_seed(0)
rnd = lambda *s: torch.randn(*s, generator=g)
Lq, Lk = (L, L) if which == "attn_self" else ((Q, L) if which == "attn_cross" else (Q, Q))
q = (rnd(B * Lq, d) * math.log2(math.e) / math.sqrt(32)).to(torch.bfloat16).to(dev)
k = rnd(B * Lk, d).to(torch.bfloat16).to(dev)
pitch = (Lk + 7) // 8 * 8
vt = torch.zeros(B * d, pitch, dtype=torch.bfloat16)
vt[:, :Lk] = rnd(B * d, Lk).to(torch.bfloat16)
vt = vt.to(dev)
mask = torch.ones(B, Lk, device=dev) if which == "attn_cross" else None
for _ in range(3):
    ops.attention(q, k, vt, B, H, Lq, Lk, key_mask=mask)
torch.cuda.synchronize()
lib = _lib.get_lib()
buf = np.zeros((4, 64, 8), dtype=np.int64)
rc = lib.svol_debug_attn_trace(C.c_void_p(buf.ctypes.data))
assert rc == 0, rc
n = (Lk + 127) // 128
t0 = buf[buf > 0].min()
names = {0: "softmax A (warp 0)", 1: "softmax B (warp 4)", 2: "MMA issuer"}
slots = {0: ["top", "s_full", "S->reg", "max", "turn", "exp", "o_full", "P stored"],
         2: ["top", "kv_full", "s_freeA", "s_freeB", "p_rdyA(j-1)", "p_rdyB(j-1)", "-", "-"]}
for role in (0, 1, 2):
    print(names[role])
    sl = slots[2 if role == 2 else 0]
    print("  j " + " ".join(f"{s:>11}" for s in sl))
    for j in range(n + (1 if role == 2 else 0)):
        row = buf[role, j]
        print(f" {j:2d} " + " ".join(f"{(v - t0) if v > 0 else -1:11d}" for v in row))
    if role < 2:
        per = np.diff(buf[role, 1:n, 7]).mean()
        d_ = buf[role, 2:n, :].astype(np.float64)
        print(f"  steady period {per:.0f} clk; mean phase lengths: "
              + ", ".join(f"{sl[i]}->{sl[i + 1]} {np.mean(d_[:, i + 1] - d_[:, i]):.0f}" for i in range(7)))
