#!/usr/bin/env python
"""One item per CTA in BOTH kernels (<= 148 items): isolates the per-key-tile code of the persistent kernel from its
scheduling.  python tools/attn_p_time.py"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from svol_b200 import ops
dev = torch.device("cuda:0")
H, d = 8, 256
for name, B, Lq, Lk in (("112 items x 13 tiles", 2, 1568, 1568), ("104 items x 49 tiles", 1, 3200, 6272), ("800 items x 49 tiles", 4, 6272, 6272),
                        ("1792 items x 13 tiles", 32, 1568, 1568)):
    g = torch.Generator(device="cpu").manual_seed(1)
    q = (torch.randn(B * Lq, d, generator=g) * math.log2(math.e) / math.sqrt(32)).to(torch.bfloat16).to(dev)
    k = torch.randn(B * Lk, d, generator=g).to(torch.bfloat16).to(dev)
    pitch = (Lk + 7) // 8 * 8
    vt = torch.zeros(B * d, pitch, dtype=torch.bfloat16); vt[:, :Lk] = torch.randn(B * d, Lk, generator=g).to(torch.bfloat16); vt = vt.to(dev)
    res = {}
    for mode in ("0", "2"):
        os.environ["SVOL_ATTN_PERSISTENT"] = mode
        for _ in range(3):
            ops.attention(q, k, vt, B, H, Lq, Lk)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.attention(q, k, vt, B, H, Lq, Lk)
        e1.record(); torch.cuda.synchronize()
        res[mode] = e0.elapsed_time(e1) / 20 * 1e3
    print(f"{name:24s} one-item kernel {res['0']:8.1f} us   persistent {res['2']:8.1f} us   ratio {res['2'] / res['0']:.3f}")
