#!/usr/bin/env python
"""One launch of each attention variant on the video self-attention shape (for ncu source-level captures)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from svol_b200 import ops
dev = torch.device("cuda:0")
B, Lq, Lk, H, d = 32, 1568, 1568, 8, 256
g = torch.Generator(device="cpu").manual_seed(0)
q = (torch.randn(B * Lq, d, generator=g) * math.log2(math.e) / math.sqrt(32)).to(torch.bfloat16).to(dev)
k = torch.randn(B * Lk, d, generator=g).to(torch.bfloat16).to(dev)
vt = torch.randn(B * d, Lk, generator=g).to(torch.bfloat16).to(dev)
for six in ("0", "1"):
    os.environ["SVOL_ATTN_SIX"] = six
    ops.attention(q, k, vt, B, H, Lq, Lk)
    torch.cuda.synchronize()
