#!/usr/bin/env python
"""How the attention kernel's CTAs pack onto the SMs (debug build: SVOL_EXTRA_NVCC_FLAGS=-DSVOL_ATTN_TRACE):
per-CTA start / end (globaltimer), duration by kind (two-tile / single-tile CTA), per-SM busy time and idle gaps.
python tools/attn_cta_timeline.py [attn_self|attn_cross]"""
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
vt = torch.randn(B * d, Lk, generator=g).to(torch.bfloat16).to(dev)
for _ in range(3):
    ops.attention(q, k, vt, B, H, Lq, Lk)
torch.cuda.synchronize()
buf = np.zeros((8192, 4), dtype=np.int64)
assert _lib.get_lib().svol_debug_attn_cta_timeline(C.c_void_p(buf.ctypes.data)) == 0
n_pairs = (Lq + 255) // 256
n = n_pairs * H * B
t = buf[:n]
t0 = t[:, 1].min()
start, end, sm = (t[:, 1] - t0) / 1e3, (t[:, 3] - t0) / 1e3, t[:, 0]
dur = end - start
rem = Lq - (n_pairs - 1) * 256
kind = (np.arange(n) % n_pairs == n_pairs - 1) & (rem <= 128)      # the last pair of every (sample, head) is a single-tile CTA
n_long = int((~kind).sum())
print(f"{which}: {n} CTAs, kernel span {end.max():.1f} us; two-tile CTAs: {n_long}, mean {dur[~kind].mean():.2f} us (min {dur[~kind].min():.2f}, max {dur[~kind].max():.2f})"
      + (f"; single-tile CTAs: {kind.sum()}, mean {dur[kind].mean():.2f} us" if kind.any() else ""))
busy, last, first, gaps, cnt = [], [], [], [], []
for s in np.unique(sm):
    i = np.where(sm == s)[0]
    o = i[np.argsort(start[i])]
    busy.append(dur[o].sum()); last.append(end[o].max()); first.append(start[o].min()); cnt.append(len(o))
    gaps.append(np.maximum(start[o][1:] - end[o][:-1], 0).sum())
busy, last, first, gaps, cnt = map(np.array, (busy, last, first, gaps, cnt))
print(f"SMs used {len(busy)}; CTAs per SM {cnt.min()}..{cnt.max()}; per-SM busy {busy.mean():.1f} us (min {busy.min():.1f}, max {busy.max():.1f}); "
      f"first start {first.mean():.2f} (max {first.max():.2f}); last end mean {last.mean():.1f} min {last.min():.1f} max {last.max():.1f}; gaps between CTAs per SM {gaps.mean():.2f} us")
# duration vs start time (does a CTA run faster when the machine empties?)
for lo in range(0, int(end.max()) + 1, 20):
    m = (start >= lo) & (start < lo + 20)
    if m.any():
        print(f"  CTAs starting in [{lo:3d}, {lo + 20:3d}) us: {m.sum():4d}, two-tile mean {dur[m & ~kind].mean() if (m & ~kind).any() else float('nan'):6.2f} us, single-tile mean {dur[m & kind].mean() if (m & kind).any() else float('nan'):6.2f} us")
