#!/usr/bin/env python
"""How much of the q|k|v projection launch is the transposed V store?  Three GEMMs over the C2 token matrix
(M = 50176, K = 256), operands flushed out of L2 before each:  (a) N = 768, every column block through the TMA store;
(b) the split launch of the forward plan: q|k (512 columns) through TMA + V^T (256 columns) through the per-head
transposed store;  (c) V^T alone (N = 256).  Per-kernel times: run under
    ncu --metrics gpu__time_duration.sum -k regex:gemm_bf16 --csv python tools/gemm_vt_probe.py"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from svol_b200 import ops

dev = torch.device("cuda:0")
B, L, d = 32, 1568, 256
M = B * L
g = torch.Generator(device="cpu").manual_seed(0)
rnd = lambda *s: torch.randn(*s, generator=g)
A = rnd(M, d).to(torch.bfloat16).to(dev)
A2 = rnd(M, d).to(torch.bfloat16).to(dev)
W3 = (rnd(3 * d, d) / math.sqrt(d)).to(torch.bfloat16).to(dev)
b3 = rnd(3 * d).to(dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for rep in range(3):
    flush.zero_()
    ops.gemm(A, W3, b3)                                               # (a) 768 columns, plain
    flush.zero_()
    ops.gemm(A, W3, b3, vt_len=L, A2=A2, split_block=2)               # (b) q|k + V^T
    flush.zero_()
    ops.gemm(A2, W3[2 * d:], b3[2 * d:], want_out=False, vt_len=L)    # (c) V^T alone
    flush.zero_()
    ops.gemm(A2, W3[2 * d:], b3[2 * d:])                              # (d) 256 columns, plain
torch.cuda.synchronize()
print("done")
