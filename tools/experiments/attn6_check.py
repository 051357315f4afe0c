#!/usr/bin/env python
"""Six-warpgroup attention variant (SVOL_ATTN_SIX=1) against the shipped kernel and an fp32 torch softmax on the same random
operands (different key-tile sizes -> not bit-identical; both must sit within bf16 rounding of the fp32 result), then the
timing of both.  python tools/attn6_check.py [reps]"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from svol_b200 import ops

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = torch.device("cuda:0")
H, d = 8, 256
LOG2E = math.log2(math.e)
cases = (("tiny", 2, 70, 45, True, 1.0), ("q", 32, 320, 320, False, 1.0), ("cross", 32, 320, 1568, True, 1.0),
         ("self", 32, 1568, 1568, False, 1.0), ("self-hot", 32, 1568, 1568, False, 4.0), ("ragged", 3, 257, 97, True, 1.0),
         ("one-tile", 2, 128, 96, False, 1.0), ("33 keys", 2, 130, 33, False, 1.0), ("long", 4, 6272, 6272, False, 1.0))
for name, B, Lq, Lk, masked, scale in cases:
    g = torch.Generator(device="cpu").manual_seed(1)
    q = (torch.randn(B * Lq, d, generator=g) * scale * LOG2E / math.sqrt(32)).to(torch.bfloat16).to(dev)
    k = torch.randn(B * Lk, d, generator=g).to(torch.bfloat16).to(dev)
    pitch = (Lk + 7) // 8 * 8
    vt = torch.zeros(B * d, pitch, dtype=torch.bfloat16)
    vt[:, :Lk] = torch.randn(B * d, Lk, generator=g).to(torch.bfloat16)
    vt = vt.to(dev)
    mask = None
    if masked:
        mask = torch.ones(B, Lk)
        for b in range(0, B, 3):
            mask[b, Lk - min(Lk - 1, 49 * (1 + b % 8)):] = 0
        mask = mask.to(dev)
    # fp32 reference: softmax in base 2 (q carries log2 e / sqrt(dh))
    qf = q.float().view(B, Lq, H, 32).permute(0, 2, 1, 3)
    kf = k.float().view(B, Lk, H, 32).permute(0, 2, 1, 3)
    vf = vt.float().view(B, H, 32, pitch)[..., :Lk].transpose(-1, -2)
    if B * H * Lq * Lk <= 2 ** 31:
        s = qf @ kf.transpose(-1, -2)
        if mask is not None:
            s = s.masked_fill(mask[:, None, None, :] == 0, float("-inf"))
        p = torch.softmax(s * math.log(2.0), dim=-1)
        ref = (p @ vf).permute(0, 2, 1, 3).reshape(B * Lq, d)
    else:
        ref = None
    os.environ["SVOL_ATTN_SIX"] = "0"
    v1 = ops.attention(q, k, vt, B, H, Lq, Lk, key_mask=mask).clone()
    torch.cuda.synchronize()
    os.environ["SVOL_ATTN_SIX"] = "1"
    outs = [ops.attention(q, k, vt, B, H, Lq, Lk, key_mask=mask).clone() for _ in range(reps)]
    torch.cuda.synchronize()
    same = sum(torch.equal(o, outs[0]) for o in outs)
    d61 = float((outs[0].float() - v1.float()).abs().max())
    msg = f"{name:9s} B={B} Lq={Lq} Lk={Lk}: six vs shipped max |diff| {d61:.4g}; {same}/{reps} repeat-identical"
    if ref is not None:
        msg += f"; vs fp32: six {float((outs[0].float() - ref).abs().max()):.4g}  shipped {float((v1.float() - ref).abs().max()):.4g}"
    print(msg, flush=True)
    res = {}
    for mode in ("0", "1"):
        os.environ["SVOL_ATTN_SIX"] = mode
        for _ in range(3):
            ops.attention(q, k, vt, B, H, Lq, Lk, key_mask=mask)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.attention(q, k, vt, B, H, Lq, Lk, key_mask=mask)
        e1.record(); torch.cuda.synchronize()
        res[mode] = e0.elapsed_time(e1) / 20 * 1e3
    print(f"          shipped {res['0']:8.1f} us   six {res['1']:8.1f} us   ratio {res['1'] / res['0']:.3f}", flush=True)
