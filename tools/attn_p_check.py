#!/usr/bin/env python
"""Persistent attention kernel against the one-item-per-CTA kernel on the same random operands (bit-identical expected:
same tiles, same order of operations), repeated; reports the first mismatch / failure.  python tools/attn_p_check.py [reps]"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from svol_b200 import ops

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device("cuda:0")
B, H, d = 32, 8, 256
for name, Lq, Lk, masked, scale in (("self", 1568, 1568, False, 1.0), ("self-hot", 1568, 1568, False, 4.0), ("cross", 320, 1568, True, 1.0),
                                   ("q", 320, 320, False, 1.0), ("long", 6272, 6272, False, 1.0)):
    Bc = 4 if name == "long" else B
    g = torch.Generator(device="cpu").manual_seed(1)
    q = (torch.randn(Bc * Lq, d, generator=g) * scale * math.log2(math.e) / math.sqrt(32)).to(torch.bfloat16).to(dev)
    k = torch.randn(Bc * Lk, d, generator=g).to(torch.bfloat16).to(dev)
    pitch = (Lk + 7) // 8 * 8
    vt = torch.zeros(Bc * d, pitch, dtype=torch.bfloat16)
    vt[:, :Lk] = torch.randn(Bc * d, Lk, generator=g).to(torch.bfloat16)
    vt = vt.to(dev)
    mask = None
    if masked:
        mask = torch.ones(Bc, Lk)
        for b in range(0, Bc, 3):
            mask[b, Lk - 49 * (1 + b % 8):] = 0
        mask = mask.to(dev)
    os.environ["SVOL_ATTN_PERSISTENT"] = "0"
    ref = ops.attention(q, k, vt, Bc, H, Lq, Lk, key_mask=mask).clone()
    torch.cuda.synchronize()
    os.environ["SVOL_ATTN_PERSISTENT"] = "1"
    bad = 0
    for r in range(reps):
        out = ops.attention(q, k, vt, Bc, H, Lq, Lk, key_mask=mask)
        torch.cuda.synchronize()
        if not torch.equal(out, ref):
            bad += 1
            diff = (out.float() - ref.float()).abs()
            print(f"  {name} rep {r}: max diff {float(diff.max()):.4g}, {int((diff > 0).sum())} elements differ, "
                  f"first bad row {int((diff.amax(1) > 0).nonzero()[0])}")
    print(f"{name}: Lq={Lq} Lk={Lk} B={Bc}: {reps - bad}/{reps} launches bit-identical to the one-item kernel")
    # back to back, no synchronisation in between (the way the step's graph replays them)
    outs = [ops.attention(q, k, vt, Bc, H, Lq, Lk, key_mask=mask) for _ in range(reps)]
    torch.cuda.synchronize()
    print(f"{name}: back-to-back: {sum(torch.equal(o, ref) for o in outs)}/{reps} identical")
