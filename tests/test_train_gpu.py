"""GPU parity of the TRAINING step (train.py:216-232): every backward kernel against a plain PyTorch fp32
reference of the same op (torch.autograd on the bf16-rounded operands), the whole head backward against the CPU
oracle's autograd (oracle/torch_port.py) and against the golden gradient samples recorded from the reference itself
(tests/golden/grads_*.npz).

Tolerances (bf16 path, stated per test): activations and activation gradients are stored in bf16 (relative 2^-8 per
element), accumulation is fp32.  Gradients are compared per tensor by relative L2 error.
"""
import ctypes as C
import math
import os
from dataclasses import replace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _rel(got: torch.Tensor, ref: torch.Tensor) -> float:
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    return float((got - ref).norm() / (ref.norm() + 1e-30))


def _bf(x):
    return x.to(torch.bfloat16)


# ------------------------------------------------------------------------------------------ row-wise kernels
def test_layernorm_bf16_forward():
    from svol_b200 import ops
    g = torch.Generator().manual_seed(0)
    z = _bf(torch.randn(1000, 256, generator=g) * 2 + 0.3).to(DEV)
    w, b = (1 + 0.1 * torch.randn(256, generator=g)).to(DEV), (0.1 * torch.randn(256, generator=g)).to(DEV)
    pos = torch.randn(320, 256, generator=g).to(DEV)
    y, yp = ops.layernorm_bf16(z, w, b, pos=pos, pos_mod=320)
    ref = torch.nn.functional.layer_norm(z.float(), (256,), w, b)
    assert _rel(y.float(), ref) < 4e-3
    assert _rel(yp.float(), ref + pos[torch.arange(1000, device=DEV) % 320]) < 4e-3


@pytest.mark.parametrize("cols,f32,with_att,n_dy", [(256, False, False, 1), (256, False, True, 2), (256, False, False, 3),
                                                     (512, True, False, 1), (768, True, False, 1), (512, False, False, 2)])
def test_layernorm_backward(cols, f32, with_att, n_dy):
    from svol_b200 import ops
    g = torch.Generator().manual_seed(cols + n_dy)
    rows = 3000
    # LayerNorm is invariant to a per-row scale up to its eps: with unit-variance rows datt is pure cancellation noise
    # (the reference's sketch gate only acts through eps), so the gated case uses rows whose variance is ~ eps
    z = torch.randn(rows, cols, generator=g) * (0.004 if with_att else 1.5) + (0.001 if with_att else 0.2)
    z = z if f32 else _bf(z).float()
    gamma = 1 + 0.1 * torch.randn(cols, generator=g)
    beta = torch.zeros(cols)
    dys = [_bf(torch.randn(rows, cols, generator=g) * 0.01) for _ in range(n_dy)]
    att = torch.rand(rows, generator=g) * 0.5 if with_att else None
    zr = z.clone().to(DEV).requires_grad_(True)
    gr, br = gamma.clone().to(DEV).requires_grad_(True), beta.to(DEV).requires_grad_(True)
    ar = att.clone().to(DEV).requires_grad_(True) if with_att else None
    zin = zr * (1 + ar[:, None]) if with_att else zr
    y = torch.nn.functional.layer_norm(zin, (cols,), gr, br)
    dy = sum(d.float() for d in dys).to(DEV)
    y.backward(dy)
    dx, datt, dg, db = ops.layernorm_backward(z.to(DEV) if f32 else _bf(z).to(DEV), [d.to(DEV) for d in dys], gamma.to(DEV),
                                              att=att.to(DEV) if with_att else None)
    assert _rel(dx.float(), zr.grad) < 6e-3, "dx"
    assert _rel(dg, gr.grad) < 2e-3 and _rel(db, br.grad) < 2e-3, "dgamma / dbeta"
    if with_att:
        assert _rel(datt, ar.grad) < 5e-3, "datt"


def test_gelu_and_activation_backward():
    from svol_b200 import ops
    g = torch.Generator().manual_seed(3)
    x = _bf(torch.randn(512, 2048, generator=g) * 2).to(DEV)
    dy = _bf(torch.randn(512, 2048, generator=g)).to(DEV)
    xr = x.float().requires_grad_(True)
    y = torch.nn.functional.gelu(xr)
    y.backward(dy.float())
    assert float((ops.gelu_bf16(x).float() - y).abs().max()) < 2e-2
    assert _rel(ops.act_backward(dy, x, ops.ACT_GELU).float(), xr.grad) < 4e-3
    h = torch.relu(x)
    assert _rel(ops.act_backward(dy, h, ops.ACT_RELU).float(), dy.float() * (h > 0)) < 1e-6


def test_transpose_colsum():
    from svol_b200 import ops
    g = torch.Generator().manual_seed(4)
    x = _bf(torch.randn(1000, 320, generator=g)).to(DEV)
    xt, cs = ops.transpose_bf16(x, want_colsum=True)
    assert xt.shape == (320, 1024)
    assert torch.equal(xt[:, :1000], x.t()) and float(xt[:, 1000:].abs().max()) == 0.0
    assert _rel(cs, x.float().sum(0)) < 1e-5
    # strided input (a column block of a wider matrix)
    wide = _bf(torch.randn(640, 512, generator=g)).to(DEV)
    xt2, _ = ops.transpose_bf16(wide[:, 256:])
    assert torch.equal(xt2[:, :640], wide[:, 256:].t())


def test_colsum():
    from svol_b200 import _lib
    g = torch.Generator().manual_seed(14)
    for rows, cols, ld in ((50176, 256, 256), (10240, 2048, 2048), (1000, 256, 512), (7, 512, 512)):
        x = _bf(torch.randn(rows, ld, generator=g)).to(DEV)
        out = torch.ones(cols, device=DEV)
        _lib.check(_lib.get_lib().svol_colsum_bf16(x.data_ptr(), ld, rows, cols, out.data_ptr(), _lib.stream_ptr()), "colsum")
        assert _rel(out, 1 + x[:, :cols].float().sum(0)) < 1e-5, (rows, cols)


def test_wgrad_through_gemm():
    """dW = dY^T X as svol_gemm_bf16 over the transposed operands (contraction over 10240 token rows)."""
    from svol_b200 import ops
    g = torch.Generator().manual_seed(5)
    rows = 10240
    dY = _bf(torch.randn(rows, 512, generator=g) * 0.05).to(DEV)
    X = _bf(torch.randn(rows, 256, generator=g)).to(DEV)
    dYt, db = ops.transpose_bf16(dY, want_colsum=True)
    Xt, _ = ops.transpose_bf16(X)
    dW = ops.gemm(dYt, Xt)["out"]
    ref = dY.float().t() @ X.float()
    assert _rel(dW.float(), ref) < 4e-3
    assert _rel(db, dY.float().sum(0)) < 1e-4


# ------------------------------------------------------------------------------------------ attention backward
def _head_t(x, B, L, Lp):
    """[B*L, 256] -> per-head transposed [B*256, Lp] (what the projection GEMM's out_vt epilogue writes)."""
    out = torch.zeros(B * 256, Lp, device=x.device, dtype=x.dtype)
    out.view(B, 256, Lp)[:, :, :L] = x.view(B, L, 256).transpose(1, 2)
    return out


@pytest.mark.parametrize("B,Lq,Lk,masked", [(2, 128, 128, False), (2, 320, 320, False), (3, 320, 1568, True),
                                            (2, 1568, 1568, False), (1, 100, 200, True),
                                            (1, 1280, 6272, True), (1, 6272, 6272, False)])      # long clip (config C4)
def test_attention_backward(B, Lq, Lk, masked):
    from svol_b200 import ops
    H, dh = 8, 32
    g = torch.Generator().manual_seed(Lq + Lk)
    c = math.log2(math.e) / math.sqrt(dh)
    q_raw = torch.randn(B * Lq, 256, generator=g)
    k = _bf(torch.randn(B * Lk, 256, generator=g)).to(DEV)
    v = _bf(torch.randn(B * Lk, 256, generator=g)).to(DEV)
    qs = _bf(q_raw * c).to(DEV)                                   # stored pre-scaled, as the projection writes it
    d_o = _bf(torch.randn(B * Lq, 256, generator=g) * 0.1).to(DEV)
    mask = None
    if masked:
        mask = torch.ones(B, Lk)
        mask[0, Lk - 37:] = 0
        mask = mask.to(DEV)
    Lqp, Lkp = (Lq + 7) // 8 * 8, (Lk + 7) // 8 * 8
    o, lse = ops.attention_train(qs, k, _head_t(v, B, Lk, Lkp), B, H, Lq, Lk, key_mask=mask)
    dq, dk, dv = ops.attention_backward(qs, k, v, _head_t(k, B, Lk, Lkp), _head_t(qs, B, Lq, Lqp), o, d_o,
                                        _head_t(d_o, B, Lq, Lqp), lse, B, H, Lq, Lk, key_mask=mask)
    torch.cuda.synchronize()
    # fp32 reference on the same (bf16-rounded) operands: q = qs / c is the unscaled query projection
    qr = (qs.float() / c).requires_grad_(True)
    kr, vr = k.float().requires_grad_(True), v.float().requires_grad_(True)
    qh = qr.view(B, Lq, H, dh).transpose(1, 2)
    kh, vh = kr.view(B, Lk, H, dh).transpose(1, 2), vr.view(B, Lk, H, dh).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) / math.sqrt(dh)
    if masked:
        s = s.masked_fill(mask[:, None, None, :] == 0, float("-inf"))
    p = torch.softmax(s, -1)
    oref = (p @ vh).transpose(1, 2).reshape(B * Lq, 256)
    lse_ref = torch.logsumexp(s, -1) * math.log2(math.e)
    oref.backward(d_o.float())
    assert _rel(o.float(), oref) < 8e-3, "forward output"
    assert float((lse[:, :, :Lq] - lse_ref).abs().max()) < 2e-2, "log-sum-exp"
    assert bool(torch.isinf(lse[:, :, Lq:]).all()), "lse padding must stay +inf"
    # tolerance: P and dS are rounded to bf16 before their MMAs (2^-8 each), fp32 accumulation
    assert _rel(dv.float(), vr.grad) < 1.5e-2, "dV"
    assert _rel(dk.float(), kr.grad) < 2.5e-2, "dK"
    assert _rel(dq.float(), qr.grad) < 2.5e-2, "dQ"
    if masked:
        assert float(dk.view(B, Lk, 256)[0, Lk - 37:].abs().max()) == 0.0 and float(dv.view(B, Lk, 256)[0, Lk - 37:].abs().max()) == 0.0


# ------------------------------------------------------------------------------------------ heads / gate / sketch branch
def test_heads_backward():
    from svol_b200 import _lib, ops
    g = torch.Generator().manual_seed(7)
    rows = 2 * 640
    hs = _bf(torch.randn(rows, 256, generator=g)).to(DEV)
    h2 = _bf(torch.relu(torch.randn(rows, 256, generator=g))).to(DEV)
    wc, bc = (torch.randn(2, 256, generator=g) * 0.06).to(DEV), torch.randn(2, generator=g).to(DEV)
    wb, bb = (torch.randn(4, 256, generator=g) * 0.06).to(DEV), torch.randn(4, generator=g).to(DEV)
    dl, db_ = (torch.randn(rows, 2, generator=g) * 0.01).to(DEV), (torch.randn(rows, 4, generator=g) * 0.01).to(DEV)
    logits, boxes = ops.heads(hs, h2, wc, bc, wb, bb)
    dhs, dh2 = torch.empty_like(hs), torch.empty_like(hs)
    dwc, dbc, dwb, dbb = torch.zeros_like(wc), torch.zeros_like(bc), torch.zeros_like(wb), torch.zeros_like(bb)
    P = _lib.ptr
    _lib.check(_lib.get_lib().svol_heads_backward(P(hs), P(h2), P(wc), P(wb), P(boxes), P(dl), P(db_), P(dhs), P(dh2), P(dwc), P(dbc),
                                                  P(dwb), P(dbb), rows, 256, _lib.stream_ptr()), "heads_backward")
    hsr = hs.float().requires_grad_(True)
    pre = torch.randn(rows, 256, generator=g).to(DEV)          # a pre-activation whose ReLU pattern equals h2's
    h2r = h2.float().requires_grad_(True)
    wcr, bcr, wbr, bbr = [t.clone().requires_grad_(True) for t in (wc, bc, wb, bb)]
    lo = hsr @ wcr.t() + bcr
    bo = torch.sigmoid(h2r @ wbr.t() + bbr)
    torch.autograd.backward([lo, bo], [dl, db_])
    assert _rel(dhs.float(), hsr.grad) < 5e-3
    assert _rel(dh2.float(), h2r.grad * (h2.float() > 0)) < 5e-3
    for got, ref in ((dwc, wcr.grad), (dbc, bcr.grad), (dwb, wbr.grad), (dbb, bbr.grad)):
        assert _rel(got, ref) < 1e-3


def test_gate_backward():
    """Backward of cross_modal_transformer.py:122-127 (sketch gate) incl. norm1, against torch.autograd."""
    from svol_b200 import _lib, ops
    B, L, d, H = 2, 392, 256, 8
    g = torch.Generator().manual_seed(8)
    # token rows with variance ~ eps: otherwise norm1 cancels the gate (LayerNorm is scale invariant up to eps) and
    # every gate gradient is rounding noise around zero
    x = _bf(torch.randn(B * L, d, generator=g) * 0.004).to(DEV)
    pos = torch.randn(B * L, d, generator=g).to(DEV)
    xp = _bf(x.float() + pos)
    sk = torch.randn(B, d, generator=g).to(DEV)
    in_w, in_b = (torch.randn(3 * d, d, generator=g) * 0.2).to(DEV), (torch.randn(3 * d, generator=g) * 0.05).to(DEV)
    ln_w, ln_b = (1 + 0.1 * torch.randn(d, generator=g)).to(DEV), (0.1 * torch.randn(d, generator=g)).to(DEV)
    mem, memp, att, scores = ops.gate(x, xp, sk, in_w, in_b, ln_w, ln_b, pos, B, L)
    dmem = _bf(torch.randn(B * L, d, generator=g) * 0.02).to(DEV)
    lib, P, st = _lib.get_lib(), _lib.ptr, _lib.stream_ptr()
    u = torch.empty(B, H, d, device=DEV)
    _lib.check(lib.svol_gate_vectors(P(sk), P(in_w), P(in_b), P(u), B, d, H, st), "gate_vectors")
    dx, datt, dg, db = ops.layernorm_backward(x, [dmem], ln_w, att=att.reshape(-1))
    dx_out, dscores, du = torch.empty_like(x), torch.empty(B, H, L, device=DEV), torch.zeros(B, H, d, device=DEV)
    _lib.check(lib.svol_gate_backward(P(xp), P(u), P(scores), P(datt), P(dx), P(dx_out), P(dscores), P(du), B, L, d, H, st), "gate_backward")
    dw, dbias, dsk = torch.zeros_like(in_w), torch.zeros_like(in_b), torch.zeros_like(sk)
    _lib.check(lib.svol_gate_vectors_backward(P(sk), P(in_w), P(in_b), P(du), P(dw), P(dbias), P(dsk), B, d, H, st), "gate_vectors_backward")
    # reference: x is the leaf; x + pos shares it (cross_modal_transformer.py:122-127)
    xr, skr = x.float().requires_grad_(True), sk.clone().requires_grad_(True)
    wr, br = in_w.clone().requires_grad_(True), in_b.clone().requires_grad_(True)
    gr, hr = ln_w.clone().requires_grad_(True), ln_b.clone().requires_grad_(True)
    kv = (xr + pos).view(B, L, d)
    qh = (skr @ wr[:d].t() + br[:d]).view(B, H, 1, d // H)
    kh = (kv @ wr[d:2 * d].t() + br[d:2 * d]).view(B, L, H, d // H).permute(0, 2, 1, 3)
    a = torch.softmax(qh @ kh.transpose(-1, -2) / math.sqrt(d // H), -1).mean(1).reshape(B * L, 1)
    y = torch.nn.functional.layer_norm(xr + a * xr, (d,), gr, hr)
    y.backward(dmem.float())
    assert _rel(mem.float(), y) < 6e-3
    assert _rel(dx_out.float(), xr.grad) < 1e-2, "dx"
    assert _rel(dg, gr.grad) < 3e-3 and _rel(db, hr.grad) < 3e-3
    assert _rel(dsk, skr.grad) < 3e-2, "dsketch"
    assert _rel(dw[:2 * d], wr.grad[:2 * d]) < 3e-2, "d in_proj_weight (q, k rows)"
    assert float(dw[2 * d:].abs().max()) == 0.0
    assert _rel(dbias[:d], br.grad[:d]) < 3e-2


def test_ln_linear_f32_backward():
    from svol_b200 import _lib, ops
    g = torch.Generator().manual_seed(9)
    rows, din, dout = 8, 512, 256
    x = torch.randn(rows, din, generator=g).to(DEV)
    lw, lb = (1 + 0.1 * torch.randn(din, generator=g)).to(DEV), (0.1 * torch.randn(din, generator=g)).to(DEV)
    w, b = (torch.randn(dout, din, generator=g) * 0.05).to(DEV), (torch.randn(dout, generator=g) * 0.05).to(DEV)
    dy = torch.randn(rows, dout, generator=g).to(DEV)
    for relu in (1, 0):
        y = ops.ln_linear_f32(x, lw, lb, w, b, bool(relu))
        leaves = [t.clone().requires_grad_(True) for t in (x, lw, lb, w, b)]
        yr = torch.nn.functional.linear(torch.nn.functional.layer_norm(leaves[0], (din,), leaves[1], leaves[2]), leaves[3], leaves[4])
        yr = torch.relu(yr) if relu else yr
        yr.backward(dy)
        dx = torch.empty_like(x)
        outs = [torch.zeros_like(t) for t in (lw, lb, w, b)]
        P = _lib.ptr
        _lib.check(_lib.get_lib().svol_ln_linear_f32_backward(P(x), P(lw), P(lb), P(w), P(y), P(dy), relu, P(dx), P(outs[0]), P(outs[1]),
                                                              P(outs[2]), P(outs[3]), rows, din, dout, 1e-5, 0.0, None, 0, _lib.stream_ptr()), "bwd")
        assert _rel(dx, leaves[0].grad) < 1e-4
        for got, leaf in zip(outs, leaves[1:]):
            assert _rel(got, leaf.grad) < 1e-4


def test_batch_sum_accum_adamw():
    from svol_b200 import _lib
    lib, P, st = _lib.get_lib(), _lib.ptr, _lib.stream_ptr()
    g = torch.Generator().manual_seed(10)
    gq = _bf(torch.randn(4 * 320, 256, generator=g)).to(DEV)
    acc = torch.ones(320, 256, device=DEV)
    _lib.check(lib.svol_batch_sum(P(gq), P(acc), 4 * 320, 256, 320, st), "batch_sum")
    assert _rel(acc, 1 + gq.float().view(4, 320, 256).sum(0)) < 1e-6
    src = _bf(torch.randn(1000, generator=g)).to(DEV)
    dst = torch.full((1000,), 2.0, device=DEV)
    _lib.check(lib.svol_accum_bf16(P(src), P(dst), 1000, 0.5, 1, st), "accum")
    assert _rel(dst, 2 + 0.5 * src.float()) < 1e-6
    # AdamW against torch.optim.AdamW, three steps
    p0 = torch.randn(5000, generator=g).to(DEV)
    p = p0.clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref], lr=1e-3, weight_decay=1e-2)
    for step in range(1, 4):
        grad = torch.randn(5000, generator=g).to(DEV)
        ref.grad = grad.clone()
        opt.step()
        _lib.check(lib.svol_adamw(P(p), P(grad), P(m), P(v), 5000, 1e-3, 0.9, 0.999, 1e-8, 1e-2, step, 1.0, st), "adamw")
    assert float((p - ref.detach()).abs().max()) < 1e-5


# ------------------------------------------------------------------------------------------ whole head
def _build(cfg, seed):
    from svol_b200 import synth
    from svol_b200.modeling import build_svanet
    model = build_svanet(cfg.to_namespace())
    sd = synth.random_state_dict(cfg, seed)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    return model.to(DEV), sd


def _head_grads(cfg, batch, seed, padded=True):
    from svol_b200 import synth
    model, sd = _build(cfg, seed)
    model.train()
    inp = synth.make_inputs(cfg, batch, seed, padded=padded)
    gl, gb = synth.make_upstream_grads(cfg, batch, seed)
    t = lambda k: torch.from_numpy(inp[k]).to(DEV)
    out = model(t("src_sketch"), t("src_sketch_mask"), t("src_video"), t("src_video_mask"))
    logits = torch.stack([a["pred_logits"] for a in out["aux_outputs"]] + [out["pred_logits"]])
    boxes = torch.stack([a["pred_boxes"] for a in out["aux_outputs"]] + [out["pred_boxes"]])
    torch.autograd.backward([logits, boxes], [torch.from_numpy(gl).to(DEV), torch.from_numpy(gb).to(DEV)])
    torch.cuda.synchronize()
    grads = {k: p.grad.detach().float().cpu() for k, p in model.named_parameters() if p.grad is not None}
    return model, sd, inp, (gl, gb), grads, logits.detach().float().cpu(), boxes.detach().float().cpu()


# per-tensor relative L2 error allowed on the bf16 path (bf16 activations, activation gradients and weight-gradient
# GEMM outputs; fp32 accumulation).  Tensors whose true gradient is ~0 (e.g. key biases, which cancel in softmax) are
# compared against the scale of the largest gradient instead.
GRAD_REL = 4e-2
# Parameters behind a ReLU (box MLP hidden layers, first input projection): the bf16 forward and the fp32 reference
# disagree on the sign of ~1 % of the pre-activations that lie within the forward error of zero, and every flipped
# mask entry moves a whole gradient element: relative L2 error ~ sqrt(flipped fraction) ~ 0.1, norms agree to ~2 % (5 % allowed).
GRAD_REL_RELU = 0.2
RELU_GATED = ("bbox_embed.layers.0.", "bbox_embed.layers.1.", "input_video_proj.0.")


def _compare(grads, ref, what, relu_tol=GRAD_REL_RELU):
    scale = max(float(v.double().norm()) for v in ref.values())
    worst = ("", 0.0)
    for k, r in ref.items():
        assert k in grads, f"{what}: no gradient for {k}"
        r = r.double()
        gk = grads[k].double()
        err = float((gk - r).norm())
        rel = err / max(float(r.norm()), 1e-4 * scale)
        gated = k.startswith(RELU_GATED)
        assert rel < (relu_tol if gated else GRAD_REL), f"{what}: {k} relative L2 error {rel:.4g}"
        assert abs(float(gk.norm()) - float(r.norm())) < (5e-2 if gated else 2e-2) * max(float(r.norm()), 1e-4 * scale), f"{what}: norm of {k}"
        if not gated and rel > worst[1]:
            worst = (k, rel)
    return worst


@pytest.mark.parametrize("cfg_name,batch,seed", [("C1a", 2, 0), ("C1b", 2, 1)])
def test_head_backward_vs_oracle_and_golden(cfg_name, batch, seed, golden_dir):
    from oracle import torch_port as tp
    from svol_b200 import synth
    cfg = replace(synth.CONFIGS[cfg_name], input_dropout=0.0)
    model, sd, inp, (gl, gb), grads, logits, boxes = _head_grads(cfg, batch, seed)
    ref, ref_logits, ref_boxes = tp.head_gradients(tp.state_dict_to_torch(sd), inp["src_sketch"], inp["src_sketch_mask"],
                                                   inp["src_video"], inp["src_video_mask"], gl, gb, nheads=cfg.nheads)
    assert float((logits - ref_logits).abs().max()) < 5e-2 and float((boxes - ref_boxes).abs().max()) < 5e-3, "training forward"
    assert "class_head.weight" not in grads
    worst = _compare(grads, ref, "vs oracle autograd")
    print(f"{cfg_name}: worst per-tensor relative gradient error vs oracle {worst[1]:.4g} ({worst[0]})")
    # golden samples recorded from the reference's own autograd (tests/golden/make_golden_grads.py)
    gold = np.load(os.path.join(golden_dir, f"grads_{cfg_name}_b2.npz"))
    stride = int(gold["stride"])
    scale = max(float(gold[k]) for k in gold.files if k.startswith("head_f64/norm/"))
    for k in gold.files:
        if not k.startswith("head_f64/sample/"):
            continue
        name = k[len("head_f64/sample/"):]
        sample = grads[name].reshape(-1)[::stride].double().numpy()
        refs = gold[k].astype(np.float64)
        norm = float(gold["head_f64/norm/" + name])
        # a strided sample of n elements carries ~ norm * sqrt(n / numel) of the tensor's L2 mass
        tol = GRAD_REL_RELU if name.startswith(RELU_GATED) else GRAD_REL
        bound = tol * max(norm, 1e-4 * scale) * math.sqrt(max(len(refs), 1) / grads[name].numel()) * 3 + 1e-12
        assert float(np.linalg.norm(sample - refs)) < bound, f"golden sample of {name}"
        assert abs(float(grads[name].double().norm()) - norm) < (5e-2 if name.startswith(RELU_GATED) else 2e-2) * max(norm, 1e-4 * scale), \
            f"golden norm of {name}"


def test_training_step_end_to_end():
    """train.py:222-232 through the drop-in API: model.train(); criterion; weighted sum; backward; the gradients agree
    with the oracle's autograd where the matchings agree, and a few fused-AdamW steps reduce the loss."""
    from oracle import torch_port as tp
    from svol_b200 import synth
    from svol_b200.modeling import build_loss
    from svol_b200.optim import FusedAdamW
    cfg = replace(synth.CONFIGS["C1b"], input_dropout=0.0)
    batch, seed = 2, 2
    model, sd = _build(cfg, seed)
    model.train()
    criterion = build_loss(cfg.to_namespace()).to(DEV).train()
    inp = synth.make_inputs(cfg, batch, seed, padded=True)
    targets_np = synth.make_targets(cfg, batch, seed, frame_mask=inp["frame_mask"])
    targets = synth.targets_to_torch(targets_np)
    t = lambda k: torch.from_numpy(inp[k]).to(DEV)
    wd = criterion.weight_dict

    def step_loss():
        out = model(t("src_sketch"), t("src_sketch_mask"), t("src_video"), t("src_video_mask"))
        loss_dict = criterion(out, targets)
        return sum(loss_dict[k] * wd[k] for k in loss_dict if k in wd), loss_dict

    total, loss_dict = step_loss()
    total.backward()
    torch.cuda.synchronize()
    grads = {k: p.grad.detach().float().cpu() for k, p in model.named_parameters() if p.grad is not None}
    ref, ref_losses, ref_idx = tp.training_step_gradients(tp.state_dict_to_torch(sd), inp["src_sketch"], inp["src_sketch_mask"],
                                                          inp["src_video"], inp["src_video_mask"], targets, cfg, wd)
    same = all(np.array_equal(g[0].numpy(), r[0].numpy()) and np.array_equal(g[1].numpy(), r[1].numpy())
               for li in range(cfg.num_layers) for g, r in zip(criterion.indices(li - 1 if li else -1), ref_idx[li]))
    for k, v in ref_losses.items():
        if same and not k.startswith("class_error"):
            assert abs(float(loss_dict[k]) - v) < 3e-2 * max(1.0, abs(v)), k
    if same:
        _compare(grads, ref, "training step vs oracle autograd")
    opt = FusedAdamW(model, lr=1e-4, weight_decay=1e-4)
    first = float(total)
    for _ in range(6):
        opt.zero_grad()
        total, _ = step_loss()
        total.backward()
        opt.step()
    total, _ = step_loss()
    assert float(total) < first, f"loss did not decrease: {first} -> {float(total)}"


def test_gradients_after_fused_adamw_steps_use_the_updated_weights():
    """Round-1 ADVICE (high): FusedAdamW updates the flat parameter buffer through a raw pointer, so neither data_ptr nor
    _version of a parameter changes -- the transposed bf16 weight copies of the dgrad GEMMs must still be refreshed.
    After N large-lr steps the gradients are compared with the oracle's autograd on the UPDATED state_dict; with stale
    W^T operands every activation gradient (and so every weight gradient upstream of the heads) would be wrong."""
    from oracle import torch_port as tp
    from svol_b200 import synth
    from svol_b200.optim import FusedAdamW
    cfg = replace(synth.CONFIGS["C1b"], input_dropout=0.0)
    batch, seed = 2, 4
    model, _ = _build(cfg, seed)
    model.train()
    inp = synth.make_inputs(cfg, batch, seed, padded=True)
    t = lambda k: torch.from_numpy(inp[k]).to(DEV)
    gl, gb = synth.make_upstream_grads(cfg, batch, seed)
    tgl, tgb = torch.from_numpy(gl).to(DEV), torch.from_numpy(gb).to(DEV)
    opt = FusedAdamW(model, lr=3e-3, weight_decay=1e-2)           # large steps: W moves by ~1e-2 per element in 4 steps
    model.train_engine.publish_grads = False

    def backward_once():
        out = model(t("src_sketch"), t("src_sketch_mask"), t("src_video"), t("src_video_mask"))
        logits = torch.stack([a["pred_logits"] for a in out["aux_outputs"]] + [out["pred_logits"]])
        boxes = torch.stack([a["pred_boxes"] for a in out["aux_outputs"]] + [out["pred_boxes"]])
        torch.autograd.backward([logits, boxes], [tgl, tgb])

    w0 = model.transformer.layers[0].mlp1.fc1.weight.detach().clone()
    for _ in range(4):
        backward_once()
        opt.step(from_engine=True)
    assert float((model.transformer.layers[0].mlp1.fc1.weight - w0).abs().mean()) > 3e-3
    model.train_engine.publish_grads = True
    for p in model.parameters():
        p.grad = None
    backward_once()
    torch.cuda.synchronize()
    sd_now = {k: v.detach().float().cpu() for k, v in model.state_dict().items()}
    ref, _, _ = tp.head_gradients(sd_now, inp["src_sketch"], inp["src_sketch_mask"], inp["src_video"], inp["src_video_mask"],
                                  gl, gb, nheads=cfg.nheads)
    grads = {k: p.grad.detach().float().cpu() for k, p in model.named_parameters() if p.grad is not None}
    # The weight-gradient GEMMs accumulate their k slices with fp32 atomics, so the weights after four LARGE steps differ from
    # run to run in their last bits, and with them the set of box-MLP pre-activations that lie within the bf16 forward error
    # of zero: the ReLU-gated tensors' error was 0.13-0.22 over repeated runs of this very test (one in six above 0.2).
    # What this test is for -- stale transposed weights in the backward -- moves EVERY tensor by O(1), far outside both bounds.
    _compare(grads, ref, "gradients after 4 FusedAdamW steps vs oracle autograd on the updated weights", relu_tol=0.3)


def test_fused_adamw_is_a_torch_optimizer():
    """Round-1 ADVICE (medium): param_groups drive the step (lr schedulers, train.py:129-137), state_dict / load_state_dict
    in torch.optim.AdamW's format (train.py:149,270), parameters without a gradient are left untouched.  Checked against
    torch.optim.AdamW itself on a small module: two groups with different lr / weight decay, one parameter that never gets a
    gradient, StepLR, a checkpoint round trip in both directions."""
    from svol_b200.optim import FusedAdamW

    def make():
        torch.manual_seed(0)
        m = torch.nn.Sequential(torch.nn.Linear(13, 7), torch.nn.Linear(7, 5), torch.nn.Linear(5, 3)).to(DEV)
        return m

    def groups(m):
        return [{"params": list(m[0].parameters()), "lr": 3e-3}, {"params": list(m[1].parameters()) + list(m[2].parameters()),
                                                                    "weight_decay": 0.05}]

    ma, mb = make(), make()
    oa = torch.optim.AdamW(groups(ma), lr=1e-2, weight_decay=1e-2)
    ob = FusedAdamW(groups(mb), lr=1e-2, weight_decay=1e-2)
    assert isinstance(ob, torch.optim.Optimizer) and len(ob.param_groups) == 2 and ob.param_groups[0]["lr"] == 3e-3
    sa = torch.optim.lr_scheduler.StepLR(oa, step_size=2, gamma=0.5)
    sb = torch.optim.lr_scheduler.StepLR(ob, step_size=2, gamma=0.5)
    x = torch.randn(11, 13, device=DEV)

    def run(m, o, s, steps):
        for _ in range(steps):
            o.zero_grad()
            (m[1](m[0](x)) ** 2).mean().backward()        # m[2] never receives a gradient
            o.step()
            s.step()

    run(ma, oa, sa, 5)
    run(mb, ob, sb, 5)
    for pa, pb in zip(ma.parameters(), mb.parameters()):
        assert torch.allclose(pa, pb, rtol=2e-5, atol=2e-6), float((pa - pb).abs().max())
    assert torch.equal(mb[2].weight, make()[2].weight)                      # no gradient -> no weight decay either
    assert ob.param_groups[0]["lr"] == oa.param_groups[0]["lr"] != 3e-3     # the scheduler reached the fused optimizer
    # checkpoint round trips: fused -> torch and torch -> fused continue identically
    sd_a, sd_b = oa.state_dict(), ob.state_dict()
    assert sorted(sd_b["state"].keys()) == sorted(sd_a["state"].keys()) and set(sd_b["state"][0]) >= {"step", "exp_avg", "exp_avg_sq"}
    mc, md = make(), make()
    mc.load_state_dict(ma.state_dict()); md.load_state_dict(mb.state_dict())
    oc = torch.optim.AdamW(groups(mc), lr=1e-2, weight_decay=1e-2)
    od = FusedAdamW(groups(md), lr=1e-2, weight_decay=1e-2)
    oc.load_state_dict(sd_b)                   # torch loads the fused optimizer's checkpoint
    od.load_state_dict(sd_a)                   # and the other way round
    sc = torch.optim.lr_scheduler.StepLR(oc, step_size=2, gamma=0.5); sc.load_state_dict(sa.state_dict())
    sd_ = torch.optim.lr_scheduler.StepLR(od, step_size=2, gamma=0.5); sd_.load_state_dict(sb.state_dict())
    run(ma, oa, sa, 3); run(mc, oc, sc, 3); run(md, od, sd_, 3)
    for pa, pc, pd in zip(ma.parameters(), mc.parameters(), md.parameters()):
        assert torch.allclose(pa, pc, rtol=2e-5, atol=2e-6) and torch.allclose(pa, pd, rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("case,cfgname", [("C1b_b2", "C1b"), ("C1a_b2", "C1a")])
def test_head_input_gradients_match_reference_golden(golden_dir, case, cfgname):
    """train.py:72 optimises backbone + head: the head's backward must hand d(loss)/d(src_video), d(loss)/d(src_sketch)
    back to whatever produced the features (model.py:18-28).  Compared with the reference's own autograd
    (tests/golden/make_golden_input_grads.py): strided samples and norms of both input gradients; the gradients reach a
    leaf tensor through ``loss.backward()`` like any torch module's.  (The sketch gradient is O(1e-11): the sketch only
    enters through the gate, which LayerNorm cancels up to its eps -- checked against an absolute bound.)"""
    from svol_b200 import synth
    g = np.load(os.path.join(golden_dir, f"grads_inputs_{case}.npz"))
    cfg = replace(synth.CONFIGS[cfgname], input_dropout=0.0)
    batch, seed = int(g["batch"]), int(g["seed"])
    model, _ = _build(cfg, seed)
    model.train()
    inp = synth.make_inputs(cfg, batch, seed, padded=bool(g["padded"]))
    t = lambda k: torch.from_numpy(inp[k]).to(DEV)
    vid, sk = t("src_video").requires_grad_(True), t("src_sketch").requires_grad_(True)
    gl, gb = synth.make_upstream_grads(cfg, batch, seed)
    out = model(sk, t("src_sketch_mask"), vid, t("src_video_mask"))
    logits = torch.stack([a["pred_logits"] for a in out["aux_outputs"]] + [out["pred_logits"]])
    boxes = torch.stack([a["pred_boxes"] for a in out["aux_outputs"]] + [out["pred_boxes"]])
    torch.autograd.backward([logits, boxes], [torch.from_numpy(gl).to(DEV), torch.from_numpy(gb).to(DEV)])
    torch.cuda.synchronize()
    assert vid.grad is not None and vid.grad.shape == vid.shape and vid.grad.dtype == torch.float32
    assert sk.grad is not None and sk.grad.shape == sk.shape
    stride = int(g["stride"])
    gv = vid.grad.detach().cpu().double().numpy().ravel()
    ref_s, ref_n = g["f64/sample/src_video"].astype(np.float64), float(g["f64/norm/src_video"])
    err = np.linalg.norm(gv[::stride] - ref_s) / np.linalg.norm(ref_s)
    print(f"{case}: d/d src_video relative L2 error of the sample {err:.4g}; norm {np.linalg.norm(gv):.6g} vs {ref_n:.6g}")
    assert err < GRAD_REL_RELU                      # behind the input projection's ReLU, like input_video_proj.0.*
    assert abs(np.linalg.norm(gv) - ref_n) < 5e-2 * ref_n
    gs = sk.grad.detach().cpu().double().numpy().ravel()
    assert np.isfinite(gs).all() and np.linalg.norm(gs) < 1e-6 * max(ref_n, 1e-3)      # reference: ~1e-11 .. 1e-8
    # parameters get their gradients in the same backward, and a detached input gets none
    assert model.input_video_proj[0].net[1].weight.grad is not None


def test_wgrad_split_k_fp32():
    """svol_gemm_bf16 in weight-gradient mode: fp32 atomic accumulation with the contraction split over the SMs."""
    from svol_b200 import _lib, ops
    g = torch.Generator().manual_seed(11)
    for rows, n_out, k_in in ((50176, 256, 256), (10240, 2048, 256), (10240, 256, 2048), (640, 512, 256), (64, 256, 768)):
        dY = _bf(torch.randn(rows, n_out, generator=g) * 0.05).to(DEV)
        X = _bf(torch.randn(rows, k_in, generator=g)).to(DEV)
        dYt, _ = ops.transpose_bf16(dY)
        Xt, _ = ops.transpose_bf16(X)
        out = torch.full((n_out, k_in), 1.0, device=DEV)
        a = _lib.GemmArgs()
        a.A, a.W, a.M, a.N, a.K, a.lda, a.ldw = dYt.data_ptr(), Xt.data_ptr(), n_out, k_in, dYt.shape[1], dYt.stride(0), Xt.stride(0)
        a.out_f32, a.ld_f32 = out.data_ptr(), k_in
        _lib.check(_lib.get_lib().svol_gemm_bf16(C.byref(a), _lib.stream_ptr()), "wgrad")
        ref = 1.0 + dY.float().t() @ X.float()
        assert _rel(out, ref) < 1e-4, (rows, n_out, k_in)


# ------------------------------------------------------------------------------------------ dropout
def test_dropout_kernels_match_host_mask():
    """Train-mode Dropout after the input LayerNorms (svanet.py:168-170): the kernels' counter-based mask equals its
    host restatement (oracle/torch_port.dropout_mask) element for element; kept values are LN(x) / (1 - p); the
    LayerNorm backward replays the same mask."""
    from oracle import torch_port as tp
    from svol_b200 import ops
    g = torch.Generator().manual_seed(12)
    p, seed_v = 0.4, 123457
    seed = torch.tensor([seed_v], dtype=torch.int64, device=DEV)
    for rows, cols, site, f32 in ((3136, 512, 0, True), (3136, 256, 1, False), (64, 768, 0, True)):
        w, b = (1 + 0.1 * torch.randn(cols, generator=g)).to(DEV), (0.1 * torch.randn(cols, generator=g)).to(DEV)
        x = torch.randn(rows, cols, generator=g) * 1.3 + 0.1
        x = x.to(DEV) if f32 else _bf(x).to(DEV)
        if f32:
            y = ops.layernorm_to_bf16(x, w, b, drop_p=p, seed=seed, site=site)
        else:
            y, _ = ops.layernorm_bf16(x, w, b, drop_p=p, seed=seed, site=site)
        mask = torch.from_numpy(tp.dropout_mask(rows, cols, p, seed_v, site)).to(DEV)
        ref = torch.nn.functional.layer_norm(x.float(), (cols,), w, b)
        assert abs(float((mask > 0).float().mean()) - (1 - p)) < 5e-3
        assert torch.equal(y == 0, (mask == 0) | (_bf(ref * mask) == 0)), "kernel mask != host restatement"
        assert _rel(y.float(), ref * mask) < 4e-3
        dy = _bf(torch.randn(rows, cols, generator=g) * 0.01).to(DEV)
        xr = x.float().requires_grad_(True)
        wr = w.clone().requires_grad_(True)
        (torch.nn.functional.layer_norm(xr, (cols,), wr, b) * mask).backward(dy.float())
        dx, _, dg, db = ops.layernorm_backward(x, [dy], w, drop_p=p, seed=seed, site=site)
        assert _rel(dx.float(), xr.grad) < 6e-3 and _rel(dg, wr.grad) < 2e-3
    # sketch branch (fused LayerNorm -> Dropout -> Linear), forward and backward
    rows, din, dout, site = 8, 512, 256, 2
    x = torch.randn(rows, din, generator=g).to(DEV)
    lw, lb = (1 + 0.1 * torch.randn(din, generator=g)).to(DEV), (0.1 * torch.randn(din, generator=g)).to(DEV)
    w, b = (torch.randn(dout, din, generator=g) * 0.05).to(DEV), (torch.randn(dout, generator=g) * 0.05).to(DEV)
    dy = torch.randn(rows, dout, generator=g).to(DEV)
    y = ops.ln_linear_f32(x, lw, lb, w, b, True, drop_p=p, seed=seed, site=site)
    mask = torch.from_numpy(tp.dropout_mask(rows, din, p, seed_v, site)).to(DEV)
    leaves = [t.clone().requires_grad_(True) for t in (x, lw, lb, w, b)]
    yr = torch.relu(torch.nn.functional.linear(torch.nn.functional.layer_norm(leaves[0], (din,), leaves[1], leaves[2]) * mask, leaves[3], leaves[4]))
    yr.backward(dy)
    assert _rel(y, yr) < 1e-4
    from svol_b200 import _lib
    dx = torch.empty_like(x)
    outs = [torch.zeros_like(t) for t in (lw, lb, w, b)]
    P = _lib.ptr
    _lib.check(_lib.get_lib().svol_ln_linear_f32_backward(P(x), P(lw), P(lb), P(w), P(y), P(dy), 1, P(dx), P(outs[0]), P(outs[1]), P(outs[2]),
                                                          P(outs[3]), rows, din, dout, 1e-5, p, P(seed), site, _lib.stream_ptr()), "bwd")
    assert _rel(dx, leaves[0].grad) < 1e-4
    for got, leaf in zip(outs, leaves[1:]):
        assert _rel(got, leaf.grad) < 1e-4


def test_head_backward_with_dropout_vs_oracle():
    """The whole head in train mode with the reference's default input_dropout = 0.4 against the oracle's autograd
    with the same (host-rebuilt) masks."""
    from oracle import torch_port as tp
    from svol_b200 import synth
    cfg = synth.CONFIGS["C1b"]                     # input_dropout = 0.4 (lib/configs.py:127)
    assert cfg.input_dropout == 0.4
    model, sd, inp, (gl, gb), grads, logits, boxes = _head_grads(cfg, 2, 3)
    seed = model.train_engine.last_seed
    ref, ref_logits, ref_boxes = tp.head_gradients(tp.state_dict_to_torch(sd), inp["src_sketch"], inp["src_sketch_mask"],
                                                   inp["src_video"], inp["src_video_mask"], gl, gb, nheads=cfg.nheads,
                                                   dropout=(cfg.input_dropout, seed))
    assert float((logits - ref_logits).abs().max()) < 5e-2 and float((boxes - ref_boxes).abs().max()) < 5e-3, "training forward"
    worst = _compare(grads, ref, "dropout 0.4 vs oracle autograd")
    print(f"dropout: worst per-tensor relative gradient error vs oracle {worst[1]:.4g} ({worst[0]})")
    # a second forward draws a different mask
    t = lambda k: torch.from_numpy(inp[k]).to(DEV)
    out2 = model(t("src_sketch"), t("src_sketch_mask"), t("src_video"), t("src_video_mask"))
    assert model.train_engine.last_seed == seed + 1
    assert float((out2["pred_logits"].detach().float().cpu() - logits[-1]).abs().max()) > 1e-4


def test_wgrad_mn_major_operands():
    """Weight-gradient GEMM on UNtransposed operands (MN-major shared-memory descriptors): dW += dY^T X with dY [rows, N_out],
    X [rows, K_in] row-major, any number of rows (the tail k-block is zero-filled by TMA)."""
    from svol_b200 import _lib
    g = torch.Generator().manual_seed(13)
    for rows, n_out, k_in in ((10240, 256, 256), (50176, 512, 256), (10240, 2048, 256), (10240, 256, 2048), (1000, 256, 512), (64, 256, 768)):
        dY = _bf(torch.randn(rows, n_out, generator=g) * 0.05).to(DEV)
        X = _bf(torch.randn(rows, k_in, generator=g)).to(DEV)
        out = torch.full((n_out, k_in), 1.0, device=DEV)
        a = _lib.GemmArgs()
        a.A, a.W, a.M, a.N, a.K, a.lda, a.ldw = dY.data_ptr(), X.data_ptr(), n_out, k_in, rows, dY.stride(0), X.stride(0)
        a.out_f32, a.ld_f32, a.mn_major = out.data_ptr(), k_in, 1
        _lib.check(_lib.get_lib().svol_gemm_bf16(C.byref(a), _lib.stream_ptr()), "wgrad mn")
        ref = 1.0 + dY.float().t() @ X.float()
        assert _rel(out, ref) < 1e-4, (rows, n_out, k_in, _rel(out, ref))
    # column slices of wider matrices (dq | dk gradients share one buffer)
    wide = _bf(torch.randn(4096, 512, generator=g) * 0.05).to(DEV)
    X = _bf(torch.randn(4096, 256, generator=g)).to(DEV)
    out = torch.zeros((256, 256), device=DEV)
    a = _lib.GemmArgs()
    a.A, a.W, a.M, a.N, a.K, a.lda, a.ldw = wide[:, 256:].data_ptr(), X.data_ptr(), 256, 256, 4096, 512, 256
    a.out_f32, a.ld_f32, a.mn_major = out.data_ptr(), 256, 1
    _lib.check(_lib.get_lib().svol_gemm_bf16(C.byref(a), _lib.stream_ptr()), "wgrad mn slice")
    assert _rel(out, wide[:, 256:].float().t() @ X.float()) < 1e-4


def test_backward_of_overwritten_forward_is_refused():
    """The saved activations live in static plan buffers: differentiating an older forward after a newer one must fail
    loudly, not return gradients of the wrong batch."""
    from svol_b200 import synth
    cfg = replace(synth.CONFIGS["C1a"], input_dropout=0.0)
    model, _ = _build(cfg, 0)
    model.train()
    inp = synth.make_inputs(cfg, 2, 0)
    t = lambda k: torch.from_numpy(inp[k]).to(DEV)
    out1 = model(t("src_sketch"), t("src_sketch_mask"), t("src_video"), t("src_video_mask"))
    out2 = model(t("src_sketch"), t("src_sketch_mask"), t("src_video"), t("src_video_mask"))
    with pytest.raises(RuntimeError, match="overwritten"):
        out1["pred_logits"].sum().backward()
    out2["pred_logits"].sum().backward()
    assert all(torch.isfinite(p.grad).all() for p in model.parameters() if p.grad is not None)
