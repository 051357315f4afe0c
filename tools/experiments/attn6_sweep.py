import math, os, sys
sys.path.insert(0, "/root/repo")
import torch
from svol_b200 import ops
dev = torch.device("cuda:0"); H, d = 8, 256
def t(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 20 * 1e3
for name, B, Lq, Lk in (("self", 32, 1568, 1568), ("cross", 32, 320, 1568), ("long", 4, 6272, 6272)):
    g = torch.Generator(device="cpu").manual_seed(1)
    q = (torch.randn(B * Lq, d, generator=g) * math.log2(math.e) / math.sqrt(32)).to(torch.bfloat16).to(dev)
    k = torch.randn(B * Lk, d, generator=g).to(torch.bfloat16).to(dev)
    pitch = (Lk + 7) // 8 * 8
    vt = torch.zeros(B * d, pitch, dtype=torch.bfloat16); vt[:, :Lk] = torch.randn(B * d, Lk, generator=g).to(torch.bfloat16); vt = vt.to(dev)
    fn = lambda: ops.attention(q, k, vt, B, H, Lq, Lk)
    os.environ["SVOL_ATTN_SIX"] = "0"
    ref = fn().clone(); base = t(fn)
    os.environ["SVOL_ATTN_SIX"] = "1"
    line = f"{name}: shipped {base:.1f} us |"
    for poll in (0, 1):
        for st in (0, 150, 280, 450):
            os.environ["SVOL_ATTN6_POLL"] = str(poll); os.environ["SVOL_ATTN6_STAGGER"] = str(st)
            o = fn(); torch.cuda.synchronize()
            dmax = float((o.float() - ref.float()).abs().max())
            line += f" p{poll}/s{st}: {t(fn):.1f}" + ("" if dmax < 0.02 else f"(BAD {dmax:.3g})")
    print(line, flush=True)
