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
buf = np.zeros((6, 64, 8), dtype=np.int64)
rc = lib.svol_debug_attn_trace(C.c_void_p(buf.ctypes.data))
assert rc == 0, rc
n = (Lk + 127) // 128
t0 = buf[buf > 0].min()
names = {0: "softmax g0 (tile 0, keys lo)", 1: "softmax g1 (tile 0, keys hi)", 2: "softmax g2 (tile 1, keys lo)",
         3: "softmax g3 (tile 1, keys hi)", 4: "MMA issuer tile 0", 5: "MMA issuer tile 1"}
slots = {0: ["top", "s_full", "S->reg", "max", "s_free arr", "exp", "o_full", "P stored"],
         4: ["QK lo", "QK hi", "PV lo", "PV hi", "-", "-", "-", "-"]}
for role in range(6):
    print(names[role])
    sl = slots[4 if role >= 4 else 0]
    print("  j " + " ".join(f"{s:>11}" for s in sl))
    for j in range(n + (1 if role >= 4 else 0)):
        row = buf[role, j]
        print(f" {j:2d} " + " ".join(f"{(v - t0) if v > 0 else -1:11d}" for v in row))
    if role < 4:
        per = np.diff(buf[role, 1:n, 7]).mean()
        d_ = buf[role, 2:n, :].astype(np.float64)
        pairs = [(0, 1), (1, 2), (2, 4), (4, 3), (3, 5), (5, 6), (6, 7)]
        print(f"  steady period {per:.0f} clk; mean phase lengths: "
              + ", ".join(f"{sl[a]}->{sl[b_]} {np.mean(d_[:, b_] - d_[:, a]):.0f}" for a, b_ in pairs))
