#!/usr/bin/env python
"""Runs one kernel of the hot path at the headline (C2) shapes a few times -- the target command for
`ncu --set full` captures:   python tools/run_kernel.py attn_self | attn_cross | attn_q | ffn_video | ffn_query | gemm_ffn_up | gemm_ffn_down | gemm_qk | gate_fused | gate_split"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from svol_b200 import ops

which = sys.argv[1] if len(sys.argv) > 1 else "attn_self"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
B, L, Q, H, d, ff = 32, 1568, 320, 8, 256, 2048
if os.environ.get("SVOL_CONFIG") == "C4":      # long clip: T=128 -> L=6272, Q=1280
    B, L, Q = 8, 6272, 1280
g = torch.Generator(device="cpu").manual_seed(0)
rnd = lambda *s: torch.randn(*s, generator=g)

if which.startswith("attn_bwd"):
    # attention backward (delta + dQ + dK/dV kernels) at the video self-attention / cross-attention shapes
    Lq, Lk = (L, L) if which == "attn_bwd_self" else (Q, L)
    Lqp, Lkp = (Lq + 7) // 8 * 8, (Lk + 7) // 8 * 8

    def head_t(x, n, npad):
        out = torch.zeros(B * d, npad, dtype=torch.bfloat16, device=dev)
        out.view(B, d, npad)[:, :, :n] = x.view(B, n, d).transpose(1, 2)
        return out
    q = (rnd(B * Lq, d) * math.log2(math.e) / math.sqrt(32)).to(torch.bfloat16).to(dev)
    k, v = rnd(B * Lk, d).to(torch.bfloat16).to(dev), rnd(B * Lk, d).to(torch.bfloat16).to(dev)
    d_o = (rnd(B * Lq, d) * 0.1).to(torch.bfloat16).to(dev)
    o, lse = ops.attention_train(q, k, head_t(v, Lk, Lkp), B, H, Lq, Lk)
    kt, qt, dot = head_t(k, Lk, Lkp), head_t(q, Lq, Lqp), head_t(d_o, Lq, Lqp)
    fn = lambda: ops.attention_backward(q, k, v, kt, qt, o, d_o, dot, lse, B, H, Lq, Lk)
elif which.startswith("attn"):
    Lq, Lk = (L, L) if which == "attn_self" else ((Q, L) if which == "attn_cross" else (Q, Q))
    q = (rnd(B * Lq, d) * math.log2(math.e) / math.sqrt(32)).to(torch.bfloat16).to(dev)
    k = rnd(B * Lk, d).to(torch.bfloat16).to(dev)
    pitch = (Lk + 7) // 8 * 8
    vt = torch.zeros(B * d, pitch, dtype=torch.bfloat16)
    vt[:, :Lk] = rnd(B * d, Lk).to(torch.bfloat16)
    vt = vt.to(dev)
    mask = torch.ones(B, Lk, device=dev) if which == "attn_cross" else None
    fn = lambda: ops.attention(q, k, vt, B, H, Lq, Lk, key_mask=mask)
elif which.startswith("gate"):
    # gate_fused (one cluster launch) | gate_split (gate_scores + gate_apply_theta); gate_vectors runs in both
    M = B * L
    x = rnd(M, d).to(torch.bfloat16).to(dev)
    theta = ops.posenc_theta(torch.ones(B, L, device=dev)).reshape(-1)
    pos = torch.from_numpy(__import__("numpy").zeros((1,), "float32"))
    sk, in_w, in_b = rnd(B, d).to(dev), (rnd(3 * d, d) * 0.08).to(dev), (rnd(3 * d) * 0.05).to(dev)
    lw, lb = torch.ones(d, device=dev), torch.zeros(d, device=dev)
    from svol_b200 import _lib
    lib, P_, st = _lib.get_lib(), _lib.ptr, _lib.stream_ptr()
    u = torch.empty(B, H, d, device=dev)
    _lib.check(lib.svol_gate_vectors(P_(sk), P_(in_w), P_(in_b), P_(u), B, d, H, st), "gate_vectors")
    scores = torch.empty(B, H, L, device=dev)
    mem, memp = torch.empty_like(x), torch.empty_like(x)
    if which == "gate_fused":       # the kernel alone (gate_vectors ran once above)
        fn = lambda: _lib.check(lib.svol_gate_fused(P_(x), P_(u), P_(lw), P_(lb), P_(theta), P_(mem), P_(memp), None, None,
                                                    B, L, d, H, 1e-5, st), "gate_fused")
    else:
        xpos = (x.float() + rnd(M, d).to(dev)).to(torch.bfloat16)

        def fn():
            _lib.check(lib.svol_gate_scores(P_(xpos), P_(u), P_(scores), B, L, d, H, st), "gate_scores")
            _lib.check(lib.svol_gate_apply_theta(P_(x), P_(scores), P_(lw), P_(lb), P_(theta), P_(mem), P_(memp), None,
                                                 B, L, d, H, 1e-5, st), "gate_apply")
elif which.startswith("ffn"):
    M = B * L if which == "ffn_video" else B * Q
    x = rnd(M, d).to(torch.bfloat16).to(dev)
    w1 = (rnd(ff, d) / math.sqrt(d)).to(torch.bfloat16).to(dev)
    w2 = (rnd(d, ff) / math.sqrt(ff)).to(torch.bfloat16).to(dev)
    b1, b2 = rnd(ff).to(dev), rnd(d).to(dev)
    ln = (torch.ones(d, device=dev), torch.zeros(d, device=dev))
    if which == "ffn_video":
        theta = ops.posenc_theta(torch.ones(B, L, device=dev)).reshape(-1)
        fn = lambda: ops.ffn(x, w1, b1, w2, b2, ln, pos_theta=theta)
    else:
        pos = rnd(Q, d).to(dev)
        fn = lambda: ops.ffn(x, w1, b1, w2, b2, ln, pos=pos, pos_mod=Q)
else:
    M = B * L
    if which == "gemm_ffn_up":
        K, N, kw = d, ff, dict(act=ops.ACT_GELU)
    elif which == "gemm_ffn_down":
        K, N, kw = ff, d, dict(residual=True, ln=True, pos=True)
    else:
        K, N, kw = d, 2 * d, dict()
    A = rnd(M, K).to(torch.bfloat16).to(dev)
    W = (rnd(N, K) / math.sqrt(K)).to(torch.bfloat16).to(dev)
    bias = rnd(N).to(dev)
    res = rnd(M, N).to(torch.bfloat16).to(dev) if kw.pop("residual", False) else None
    ln = (torch.ones(N, device=dev), torch.zeros(N, device=dev)) if kw.pop("ln", False) else None
    pos = rnd(M, N).to(dev) if kw.pop("pos", False) else None
    fn = lambda: ops.gemm(A, W, bias, residual=res, ln=ln, pos=pos, **kw)

for _ in range(reps):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    fn()
e1.record()
torch.cuda.synchronize()
print(f"{which}: {e0.elapsed_time(e1) / reps * 1e3:.1f} us per launch")
