#!/usr/bin/env python
"""Phase timeline of the GEMM kernel's epilogue (debug build only):
    touch svol_b200/csrc/gemm_tc.cu; SVOL_EXTRA_NVCC_FLAGS=-DSVOL_GEMM_TRACE bash svol_b200/csrc/build.sh
    python tools/gemm_trace.py [gemm_ffn_up|gemm_ffn_down|gemm_qk]"""
import ctypes as C
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from svol_b200 import _lib, ops

which = sys.argv[1] if len(sys.argv) > 1 else "gemm_ffn_up"
dev = torch.device("cuda:0")
B, L, d, ff = 32, 1568, 256, 2048
g = torch.Generator(device="cpu").manual_seed(0)
rnd = lambda *s: torch.randn(*s, generator=g)
M = B * L
if which == "gemm_ffn_up":
    K, N, kw = d, ff, dict(act=ops.ACT_GELU)
elif which == "gemm_ffn_down":
    K, N, kw = ff, d, dict(residual=True, ln=True, pos=True)
elif which == "gemm_sa_out":            # attention output projection + residual + LayerNorm (cross_modal_transformer.py:140-141)
    K, N, kw = d, d, dict(residual=True, ln=True)
else:
    K, N, kw = d, 2 * d, dict()
A = rnd(M, K).to(torch.bfloat16).to(dev)
W = (rnd(N, K) / math.sqrt(K)).to(torch.bfloat16).to(dev)
bias = rnd(N).to(dev)
res = rnd(M, N).to(torch.bfloat16).to(dev) if kw.pop("residual", False) else None
ln = (torch.ones(N, device=dev), torch.zeros(N, device=dev)) if kw.pop("ln", False) else None
pos = rnd(M, N).to(dev) if kw.pop("pos", False) else None
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    if os.environ.get("SVOL_TRACE_COLD", "1") == "1":
        flush.zero_()                     # operands come from HBM, as they do inside the step
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.gemm(A, W, bias, residual=res, ln=ln, pos=pos, **kw)
    e1.record()
torch.cuda.synchronize()
print(f"{which}: M={M} N={N} K={K}: {e0.elapsed_time(e1) * 1e3:.1f} us (last launch, events)")
buf = np.zeros((2, 32, 8), dtype=np.int64)
rc = _lib.get_lib().svol_debug_gemm_trace(C.c_void_p(buf.ctypes.data))
assert rc == 0
t0 = buf[buf > 0].min()
names = ["top", "tmem_full", "acc->reg", "bias+act", "residual", "LN", "out", "out_pos/vt"]
print("epilogue warp 4 of CTA 0:  " + " ".join(f"{n:>10}" for n in names))
n_it = int((buf[0, :, 0] > 0).sum())
for it in range(n_it):
    print(f" tile {it:2d}                 " + " ".join(f"{(v - t0) if v > 0 else -1:10d}" for v in buf[0, it]))
d_ = buf[0, 1:n_it].astype(np.float64)
print("mean phase lengths: " + ", ".join(f"{names[i]}->{names[i + 1]} {np.mean(d_[:, i + 1] - d_[:, i]):.0f}" for i in range(7))
      + f"; tile period {np.diff(buf[0, 1:n_it, 7]).mean():.0f}")
print("MMA issuer: (top, tmem_empty ok, committed)")
for it in range(min(n_it, 8)):
    print(f" tile {it:2d} " + " ".join(f"{(v - t0) if v > 0 else -1:10d}" for v in buf[1, it, :3]))
print(f"MMA tile period {np.diff(buf[1, 1:n_it, 2]).mean():.0f}; issue time (tmem_empty ok -> committed) {np.mean(buf[1, 1:n_it, 2] - buf[1, 1:n_it, 1]):.0f}")
