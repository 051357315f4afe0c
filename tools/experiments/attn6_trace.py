#!/usr/bin/env python
"""Phase timeline of the six-warpgroup attention variant (debug build: SVOL_EXTRA_NVCC_FLAGS=-DSVOL_ATTN_TRACE)."""
import ctypes as C, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from svol_b200 import _lib, ops
os.environ["SVOL_ATTN_SIX"] = "1"
dev = torch.device("cuda:0")
B, Lq, Lk, H, d = 32, 1568, 1568, 8, 256
g = torch.Generator(device="cpu").manual_seed(0)
q = (torch.randn(B * Lq, d, generator=g) * math.log2(math.e) / math.sqrt(32)).to(torch.bfloat16).to(dev)
k = torch.randn(B * Lk, d, generator=g).to(torch.bfloat16).to(dev)
vt = torch.randn(B * d, Lk, generator=g).to(torch.bfloat16).to(dev)
for _ in range(3):
    ops.attention(q, k, vt, B, H, Lq, Lk)
torch.cuda.synchronize()
buf = np.zeros((6, 64, 8), dtype=np.int64)
assert _lib.get_lib().svol_debug_attn6_trace(C.c_void_p(buf.ctypes.data)) == 0
n = (Lk + 95) // 96
t0 = buf[buf > 0].min()
sl = ["top", "s_full", "S->reg", "max", "exp", "o_full", "P stored"]
for role in range(6):
    print(f"softmax warpgroup {role} (query tile {role // 3}, part {role % 3})")
    print("  j " + " ".join(f"{s:>9}" for s in sl))
    for j in range(n):
        print(f" {j:2d} " + " ".join(f"{(v - t0) if v > 0 else -1:9d}" for v in buf[role, j, :7]))
    d_ = buf[role, 3:n - 1, :7].astype(np.float64)
    per = np.diff(buf[role, 2:n - 1, 6]).mean()
    print(f"  steady period {per:.0f} clk; phases: " + ", ".join(f"{sl[a]}->{sl[a + 1]} {np.mean(d_[:, a + 1] - d_[:, a]):.0f}" for a in range(6)))
