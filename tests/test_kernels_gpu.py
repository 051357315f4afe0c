"""GPU parity of the individual kernels, called through the C ABI (svol_b200.ops) and compared with
the CPU oracle's functions on the same seeded inputs.  bf16 tolerances are stated per test: operands
are rounded to bf16 on both sides, so what remains is fp32 accumulation order plus the bf16 rounding
of the stored result (relative 2^-8) -- and, for attention, of the probabilities."""
import os
import math

import numpy as np
import pytest
import torch

from oracle import svol_oracle as orc

pytestmark = pytest.mark.gpu

BF16_REL = 2.0 ** -8


def _dev():
    return torch.device("cuda:0")


def _bf16(x: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(x).to(torch.bfloat16)


def _f(t: torch.Tensor) -> np.ndarray:
    return t.detach().float().cpu().numpy()


def _assert_close(got, ref, rel=BF16_REL, atol=2e-3, what=""):
    err = np.abs(got - ref)
    bound = rel * np.abs(ref) + atol
    bad = err > bound
    assert not bad.any(), f"{what}: {bad.sum()} / {bad.size} outside tolerance, max err {err.max():.4g} at ref {ref.flat[err.argmax()]:.4g}"


# ------------------------------------------------------------------------------------------ GEMM
GEMM_CASES = [
    # M, N, K, flags
    (128, 256, 64, dict()),
    (300, 256, 256, dict(bias=True)),
    (1000, 512, 256, dict(bias=True)),
    (517, 2048, 256, dict(bias=True, act="gelu")),
    (640, 256, 2048, dict(bias=True, residual=True, ln=True, pos=True)),
    (392, 256, 512, dict(bias=True, act="relu", ln=True)),
    (3136, 256, 256, dict(bias=True, residual=True, ln=True, pos=True, pos_mod=320)),
    (2 * 196, 256, 256, dict(bias=True, vt_len=196)),
    (40000, 256, 256, dict(bias=True)),          # > 148 tiles: persistent loop, both TMEM stages, phase wrap
    (2 * 392, 256, 256, dict(bias=True, theta=True)),   # out_pos with the sine encoding evaluated in the epilogue
]


@pytest.mark.parametrize("plain", [False, True], ids=["tcgen05", "plain"])
@pytest.mark.parametrize("M,N,K,flags", GEMM_CASES)
def test_gemm_epilogues(M, N, K, flags, plain):
    from svol_b200 import ops
    if plain and (M > 5000 or flags.get("theta")):
        pytest.skip("plain kernel: small cases, table positions only")
    rng = np.random.RandomState(M + N + K)
    A = _bf16(rng.standard_normal((M, K)).astype(np.float32))
    W = _bf16((rng.standard_normal((N, K)) / math.sqrt(K)).astype(np.float32))
    bias = torch.from_numpy(rng.standard_normal(N).astype(np.float32)) if flags.get("bias") else None
    res_t = _bf16(rng.standard_normal((M, N)).astype(np.float32)) if flags.get("residual") else None
    ln = None
    if flags.get("ln"):
        ln = (torch.from_numpy((1 + 0.1 * rng.standard_normal(N)).astype(np.float32)),
              torch.from_numpy((0.1 * rng.standard_normal(N)).astype(np.float32)))
    pos_mod = flags.get("pos_mod", 0)
    pos = torch.from_numpy(rng.standard_normal((pos_mod or M, N)).astype(np.float32)) if flags.get("pos") else None
    act = {"relu": ops.ACT_RELU, "gelu": ops.ACT_GELU}.get(flags.get("act"), ops.ACT_NONE)
    vt_len = flags.get("vt_len", 0)

    d = _dev()
    cu = lambda t: None if t is None else t.to(d)
    theta = None
    if flags.get("theta"):
        mask = np.ones((2, M // 2), np.float32)
        mask[1, -100:] = 0
        pos = torch.from_numpy(orc.position_embedding_sine(mask.astype(bool), N).reshape(M, N).astype(np.float32))
        theta = ops.posenc_theta(torch.from_numpy(mask).to(d)).reshape(-1)
    out = ops.gemm(cu(A), cu(W), cu(bias), act=act, residual=cu(res_t), ln=None if ln is None else (cu(ln[0]), cu(ln[1])),
                   pos=None if theta is not None else cu(pos), pos_mod=pos_mod, vt_len=vt_len, plain=plain, pos_theta=theta)
    torch.cuda.synchronize()

    ref = A.float().numpy() @ W.float().numpy().T
    if bias is not None:
        ref = ref + bias.numpy()
    if flags.get("act") == "relu":
        ref = np.maximum(ref, 0)
    elif flags.get("act") == "gelu":
        ref = orc.gelu(ref.astype(np.float32))
    if res_t is not None:
        ref = ref + res_t.float().numpy()
    if ln is not None:
        ref = orc.layer_norm(ref.astype(np.float32), ln[0].numpy(), ln[1].numpy())
    _assert_close(_f(out["out"]), ref, what="out")
    if pos is not None:
        prow = np.arange(M) % pos_mod if pos_mod else np.arange(M)
        _assert_close(_f(out["out_pos"]), ref + pos.numpy()[prow], what="out_pos")
    if vt_len:
        B = M // vt_len
        vt = _f(out["out_vt"]).reshape(B, N, -1)
        want = ref.reshape(B, vt_len, N).transpose(0, 2, 1)
        _assert_close(vt[:, :, :vt_len], want, what="out_vt")
        assert (vt[:, :, vt_len:] == 0).all()


def test_gemm_l2_prefetch_changes_nothing_but_timing(monkeypatch):
    """The resident-weight GEMM's producer prefetches the next A tile and the tile's residual rows into L2 (gemm_tc.cu);
    a hint only: outputs with SVOL_GEMM_L2_PREFETCH=0 / 1 are bit-identical, also when the last row tile is ragged and a
    CTA walks several tiles."""
    from svol_b200 import ops
    M, N, K = 148 * 128 * 2 + 77, 256, 256
    rng = np.random.RandomState(11)
    d = _dev()
    A = _bf16(rng.standard_normal((M, K)).astype(np.float32)).to(d)
    W = _bf16((rng.standard_normal((N, K)) / math.sqrt(K)).astype(np.float32)).to(d)
    bias = torch.from_numpy(rng.standard_normal(N).astype(np.float32)).to(d)
    res = _bf16(rng.standard_normal((M, N)).astype(np.float32)).to(d)
    ln = (torch.ones(N, device=d), torch.zeros(N, device=d))
    outs = []
    for pf in ("0", "1"):
        monkeypatch.setenv("SVOL_GEMM_L2_PREFETCH", pf)
        outs.append(ops.gemm(A, W, bias, residual=res, ln=ln)["out"].clone())
        torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])
    ref = orc.layer_norm((A.float() @ W.float().T + bias + res.float()).cpu().numpy(), np.ones(N, np.float32), np.zeros(N, np.float32))
    _assert_close(_f(outs[1]), ref, what="out")


def test_gemm_split_launch_qk_and_v():
    """One launch: columns [0,512) = (x + pos) Wqk^T written row-major, columns [512,768) = x Wv^T written per-head
    transposed (the q/k and v projections of cross_modal_transformer.py:137-139 share a kernel)."""
    from svol_b200 import ops
    rng = np.random.RandomState(11)
    B, Lt, K = 2, 196, 256
    M = B * Lt
    x = _bf16(rng.standard_normal((M, K)).astype(np.float32))
    xp = _bf16(rng.standard_normal((M, K)).astype(np.float32))
    W = _bf16((rng.standard_normal((768, K)) / 16).astype(np.float32))
    bias = torch.from_numpy(rng.standard_normal(768).astype(np.float32))
    d = _dev()
    out = ops.gemm(xp.to(d), W.to(d), bias.to(d), vt_len=Lt, A2=x.to(d), split_block=2)
    torch.cuda.synchronize()
    ref_qk = xp.float().numpy() @ W.float().numpy()[:512].T + bias.numpy()[:512]
    ref_v = x.float().numpy() @ W.float().numpy()[512:].T + bias.numpy()[512:]
    assert out["out"].shape == (M, 512)
    _assert_close(_f(out["out"]), ref_qk, what="qk")
    vt = _f(out["out_vt"]).reshape(B, 256, -1)
    _assert_close(vt[:, :, :Lt], ref_v.reshape(B, Lt, 256).transpose(0, 2, 1), what="vt")
    assert (vt[:, :, Lt:] == 0).all()


# ------------------------------------------------------------------------------------------ attention
ATTN_CASES = [
    # B, Lq, Lk, masked
    (1, 128, 128, False),
    (2, 128, 256, False),
    (2, 320, 320, False),       # query self-attention shape
    (2, 200, 392, False),       # ragged tails on both sides
    (2, 320, 1568, True),       # cross-attention with padded keys
    (1, 1568, 1568, False),     # video self-attention shape
    (2, 96, 40, False),         # fewer than 64 keys: the upper-half softmax warpgroups never run
    (2, 130, 192, False),       # Lk % 128 == 64: upper half of the last key tile empty (skipped)
    (2, 130, 193, False),       # Lk % 128 == 65: one key in the upper half of the last tile
    (2, 64, 1000, False),       # single-tile CTA only (even / odd key tiles on the two tile slots), 8 key tiles
    (2, 100, 1408, True),       # single-tile CTA, 11 key tiles (odd count, ring wraps twice), padded keys
    (1, 384, 128, False),       # one key tile: the odd tile slot of the single-tile CTA has no work
    (2, 64, 936, False),        # single-tile CTA, 8 key tiles, the last one (odd slot) without an upper half
    (2, 320, 936, True),        # same key layout behind a full CTA + a single-tile CTA, padded keys
    (1, 40, 520, False),        # single-tile CTA, 5 key tiles: last tile on the even slot, 8 keys in it
    (1, 900, 700, False),       # three full CTAs + a single-tile CTA, ragged keys
    # single-tile CTAs with at most 64 query rows: the rows are loaded twice and the two copies split each half tile's keys
    (3, 33, 517, False),        # 33 rows; last key tile holds 5 keys (only copy 0 of the lower half has any)
    (2, 1, 640, True),          # one query row, padded keys (whole 32-key parts masked)
    (2, 64, 596, True),         # last tile: lower half full, upper half 20 keys (copy 1 of it empty), padded keys
    (1, 288, 1024, False),      # a full CTA + a 32-row duplicated tile, 8 key tiles
    # short-key kernel: ragged key chunk (keys beyond Lk masked out, V^T chunk straddling Lk), ragged last query block, Lk = 512
    (3, 70, 37, False),
    (2, 129, 457, False),
    (2, 320, 512, False),
]


@pytest.mark.parametrize("path", ["tcgen05", "short", "plain"])
@pytest.mark.parametrize("B,Lq,Lk,masked", ATTN_CASES)
def test_attention(B, Lq, Lk, masked, path, monkeypatch):
    """path: "tcgen05" = attention_tc_kernel for every shape (SVOL_ATTN_SMALL=0), "short" = the default dispatch for the
    shapes the short-key kernel takes (attn_small.cu: Lk <= 512, no mask), "plain" = the SIMT twin."""
    from svol_b200 import ops
    if path == "short" and (masked or Lk > 512):
        pytest.skip("not a short-key shape")
    monkeypatch.setenv("SVOL_ATTN_SMALL", "0" if path == "tcgen05" else "1")
    plain = path == "plain"
    H, dh = 8, 32
    rng = np.random.RandomState(Lq * 7 + Lk)
    scale = math.log2(math.e) / math.sqrt(dh)
    q = _bf16((rng.standard_normal((B * Lq, H * dh)) * scale * 1.5).astype(np.float32))
    k = _bf16(rng.standard_normal((B * Lk, H * dh)).astype(np.float32))
    v = _bf16(rng.standard_normal((B * Lk, H * dh)).astype(np.float32))
    pitch = (Lk + 7) // 8 * 8
    vt = torch.zeros((B * H * dh, pitch), dtype=torch.bfloat16)
    vt[:, :Lk] = v.view(B, Lk, H * dh).permute(0, 2, 1).reshape(B * H * dh, Lk)
    mask = None
    if masked:
        mask = np.ones((B, Lk), np.float32)
        mask[0, Lk - 300:] = 0
        mask[1, Lk - 49:] = 0
    d = _dev()
    out = ops.attention(q.to(d), k.to(d), vt.to(d), B, H, Lq, Lk,
                        key_mask=None if mask is None else torch.from_numpy(mask).to(d), plain=plain)
    torch.cuda.synchronize()
    qf = q.float().numpy().reshape(B, Lq, H, dh).transpose(0, 2, 1, 3)
    kf = k.float().numpy().reshape(B, Lk, H, dh).transpose(0, 2, 1, 3)
    vf = v.float().numpy().reshape(B, Lk, H, dh).transpose(0, 2, 1, 3)
    s = (qf @ kf.transpose(0, 1, 3, 2)) * math.log(2.0)        # kernel works in base 2
    if mask is not None:
        s = np.where(mask[:, None, None, :] != 0, s, -np.inf)
    ref = (orc.softmax(s.astype(np.float32)) @ vf).transpose(0, 2, 1, 3).reshape(B * Lq, H * dh)
    got = _f(out)
    err = np.abs(got - ref)
    # probabilities are rounded to bf16 before P V (relative 2^-9 each) and the result to bf16
    assert err.max() < 2.5e-2, f"max err {err.max()}"
    assert err.mean() < 2.5e-3, f"mean err {err.mean()}"


# ------------------------------------------------------------------------------------------ row-wise kernels
@pytest.mark.parametrize("rows,cols", [(1000, 512), (77, 768), (64, 64), (5, 1024)])
def test_layernorm_to_bf16(rows, cols):
    from svol_b200 import ops
    rng = np.random.RandomState(rows)
    x = (rng.standard_normal((rows, cols)) * 2 + 0.5).astype(np.float32)
    w = (1 + 0.1 * rng.standard_normal(cols)).astype(np.float32)
    b = (0.1 * rng.standard_normal(cols)).astype(np.float32)
    d = _dev()
    y = ops.layernorm_to_bf16(torch.from_numpy(x).to(d), torch.from_numpy(w).to(d), torch.from_numpy(b).to(d))
    _assert_close(_f(y), orc.layer_norm(x, w, b), rel=2.0 ** -8, atol=1e-5, what="layernorm")


def test_ln_linear_f32():
    from svol_b200 import ops
    rng = np.random.RandomState(3)
    x = rng.standard_normal((7, 512)).astype(np.float32)
    lw, lb = (1 + 0.1 * rng.standard_normal(512)).astype(np.float32), (0.1 * rng.standard_normal(512)).astype(np.float32)
    w, b = (rng.standard_normal((256, 512)) / 22).astype(np.float32), rng.standard_normal(256).astype(np.float32)
    d = _dev()
    t = lambda a: torch.from_numpy(a).to(d)
    for relu in (False, True):
        y = _f(ops.ln_linear_f32(t(x), t(lw), t(lb), t(w), t(b), relu))
        ref = orc.linear(orc.layer_norm(x, lw, lb), w, b)
        ref = np.maximum(ref, 0) if relu else ref
        assert np.abs(y - ref).max() < 2e-5


@pytest.mark.parametrize("B,L", [(2, 1568), (3, 32), (2, 100)])
def test_posenc_sine(B, L):
    from svol_b200 import ops
    mask = np.ones((B, L), np.float32)
    mask[0, L - L // 5:] = 0
    pos = _f(ops.posenc_sine(torch.from_numpy(mask).to(_dev()), 256))
    ref = orc.position_embedding_sine(mask != 0, 256)
    # sin/cos of arguments up to 2*pi: one fp32 ulp of the argument (4.8e-7) bounds the difference
    assert np.abs(pos - ref).max() < 2e-6


def test_add_pos_broadcast():
    from svol_b200 import ops
    rng = np.random.RandomState(0)
    x = rng.standard_normal((320, 256)).astype(np.float32)
    p = rng.standard_normal((320, 256)).astype(np.float32)
    d = _dev()
    y = _f(ops.add_pos_bf16(torch.from_numpy(x).to(d), None, 960, mod=320))
    assert np.array_equal(y, np.tile(_f(_bf16(x)), (3, 1)))
    y2 = _f(ops.add_pos_bf16(torch.from_numpy(x).to(d), torch.from_numpy(p).to(d), 640, mod=320))
    assert np.array_equal(y2, np.tile(_f(_bf16(x + p)), (2, 1)))


@pytest.mark.parametrize("theta", [False, True], ids=["table", "theta"])
@pytest.mark.parametrize("B,L", [(2, 1568), (3, 70)])
def test_gate(B, L, theta):
    """Sketch-conditioned gate vs the oracle's full multi-head attention weights (the CUDA path never
    forms K; the key bias cancels in the softmax)."""
    from svol_b200 import ops
    rng = np.random.RandomState(L)
    d_model, H = 256, 8
    x = _bf16(rng.standard_normal((B, L, d_model)).astype(np.float32))
    mask = np.ones((B, L), bool)
    mask[-1, L - L // 5:] = False                 # one padded sample: its angles are normalised by the valid length
    pos = orc.position_embedding_sine(mask, d_model)
    xpos = _bf16(x.float().numpy() + pos)
    skch = rng.standard_normal((B, d_model)).astype(np.float32)
    in_w = (rng.standard_normal((3 * d_model, d_model)) * 0.08).astype(np.float32)
    in_b = (rng.standard_normal(3 * d_model) * 0.05).astype(np.float32)
    lw, lb = (1 + 0.1 * rng.standard_normal(d_model)).astype(np.float32), (0.1 * rng.standard_normal(d_model)).astype(np.float32)
    dv = _dev()
    t = lambda a: torch.from_numpy(a).to(dv)
    pos_arg = ops.posenc_theta(t(mask.astype(np.float32))).reshape(-1) if theta else t(pos.reshape(B * L, d_model))
    mem, mem_pos, att, _ = ops.gate(x.reshape(B * L, d_model).to(dv), xpos.reshape(B * L, d_model).to(dv), t(skch), t(in_w),
                                    t(in_b), t(lw), t(lb), pos_arg, B, L, H, pos_is_theta=theta)
    xf, xpf = x.float().numpy(), xpos.float().numpy()
    _, att_ref = orc.multihead_attention(skch[:, None, :], xpf, xpf, in_w, in_b, np.eye(d_model, dtype=np.float32),
                                         np.zeros(d_model, np.float32), H)
    att_ref = att_ref[:, 0, :]
    assert np.abs(_f(att) - att_ref).max() < 1e-3 * att_ref.max() + 1e-7
    mem_ref = orc.layer_norm(xf + att_ref[..., None] * xf, lw, lb)
    _assert_close(_f(mem).reshape(B, L, d_model), mem_ref, atol=1e-3, what="mem")
    _assert_close(_f(mem_pos).reshape(B, L, d_model), mem_ref + pos, atol=2e-3, what="mem_pos")


@pytest.mark.parametrize("B,L", [(2, 1568), (3, 70), (2, 5), (1, 3384), (33, 257)])
def test_gate_fused(B, L):
    """One-launch gate (cluster per sample, rows kept in shared memory, softmax statistics exchanged through
    distributed shared memory) vs the oracle's multi-head attention weights; scores are formed from x + pos in fp32
    (the split kernels read a bf16-rounded x + pos), so the oracle gets the unrounded sum."""
    from svol_b200 import ops
    rng = np.random.RandomState(L + 1)
    d_model, H = 256, 8
    x = _bf16(rng.standard_normal((B, L, d_model)).astype(np.float32))
    mask = np.ones((B, L), bool)
    mask[-1, L - L // 5:] = False
    pos = orc.position_embedding_sine(mask, d_model)
    skch = rng.standard_normal((B, d_model)).astype(np.float32)
    in_w = (rng.standard_normal((3 * d_model, d_model)) * 0.08).astype(np.float32)
    in_b = (rng.standard_normal(3 * d_model) * 0.05).astype(np.float32)
    lw, lb = (1 + 0.1 * rng.standard_normal(d_model)).astype(np.float32), (0.1 * rng.standard_normal(d_model)).astype(np.float32)
    dv = _dev()
    t = lambda a: torch.from_numpy(a).to(dv)
    theta = ops.posenc_theta(t(mask.astype(np.float32))).reshape(-1)
    mem, mem_pos, att, scores = ops.gate_fused(x.reshape(B * L, d_model).to(dv), t(skch), t(in_w), t(in_b), t(lw), t(lb), theta,
                                               B, L, H)
    torch.cuda.synchronize()
    xf = x.float().numpy()
    xpf = (xf + pos).astype(np.float32)
    _, att_ref = orc.multihead_attention(skch[:, None, :], xpf, xpf, in_w, in_b, np.eye(d_model, dtype=np.float32),
                                         np.zeros(d_model, np.float32), H)
    att_ref = att_ref[:, 0, :]
    assert np.abs(_f(att) - att_ref).max() < 1e-3 * att_ref.max() + 1e-7
    assert np.allclose(_f(att).sum(1), 1.0, atol=1e-4)
    # the exported scores reproduce att through a plain softmax (what the training backward consumes)
    sc = _f(scores).astype(np.float64)
    p = np.exp(sc - sc.max(-1, keepdims=True))
    assert np.abs((p / p.sum(-1, keepdims=True)).mean(1) - _f(att)).max() < 1e-5
    mem_ref = orc.layer_norm(xf + att_ref[..., None] * xf, lw, lb)
    _assert_close(_f(mem).reshape(B, L, d_model), mem_ref, atol=1e-3, what="mem")
    _assert_close(_f(mem_pos).reshape(B, L, d_model), mem_ref + pos, atol=2e-3, what="mem_pos")


def test_gate_fused_rejects_long_clips():
    from svol_b200 import ops, _lib
    assert _lib.get_lib().svol_gate_fused_supported(3384) == 1
    assert _lib.get_lib().svol_gate_fused_supported(3385) == 0
    dv = _dev()
    with pytest.raises(ValueError):
        ops.gate_fused(torch.zeros(6272, 256, dtype=torch.bfloat16, device=dv), torch.zeros(1, 256, device=dv),
                       torch.zeros(768, 256, device=dv), torch.zeros(768, device=dv), torch.ones(256, device=dv),
                       torch.zeros(256, device=dv), torch.zeros(6272, device=dv), 1, 6272)


def test_heads():
    from svol_b200 import ops
    rng = np.random.RandomState(5)
    rows = 1000
    hs, h2 = _bf16(rng.standard_normal((rows, 256)).astype(np.float32)), _bf16(np.abs(rng.standard_normal((rows, 256))).astype(np.float32))
    wc, bc = (rng.standard_normal((2, 256)) / 16).astype(np.float32), rng.standard_normal(2).astype(np.float32)
    wb, bb = (rng.standard_normal((4, 256)) / 16).astype(np.float32), rng.standard_normal(4).astype(np.float32)
    dv = _dev()
    t = lambda a: torch.from_numpy(a).to(dv)
    logits, boxes = ops.heads(hs.to(dv), h2.to(dv), t(wc), t(bc), t(wb), t(bb))
    assert np.abs(_f(logits) - orc.linear(hs.float().numpy(), wc, bc)).max() < 1e-4
    assert np.abs(_f(boxes) - orc.sigmoid(orc.linear(h2.float().numpy(), wb, bb))).max() < 1e-5


# ------------------------------------------------------------------------------------------ fused FFN
FFN_CASES = [
    # M, ff, pos mode
    (128, 256, None),
    (300, 512, None),
    (640, 2048, "full"),
    (3136, 2048, "mod"),          # pos rows repeat with period 320 (query embedding)
    (20000, 2048, "full"),        # > 148 tiles: persistent loop, barrier phase wrap across tiles
    (2 * 1568, 2048, "theta"),    # sine positional encoding evaluated inside the kernel (padded second sample)
]


def test_ffn_gelu_range_and_wide_epilogue(monkeypatch):
    """The chunk epilogue's GELU is relu(t) - |t|/2 erfc(|t|/sqrt 2) with a degree-3 fit of log2 erfc (ffn_tc.cu): checked
    against the oracle's erf form over pre-activations spanning +-12 (far tails, the clamp at 4 sqrt 2, values near 0), for the
    default 8 epilogue warps and for the 16-warp instantiation (SVOL_FFN_EPI_WARPS=16)."""
    from svol_b200 import ops
    d, ff, M = 256, 512, 384
    rng = np.random.RandomState(5)
    x = _bf16(rng.standard_normal((M, d)).astype(np.float32))
    w1 = _bf16((3.0 * rng.standard_normal((ff, d)) / math.sqrt(d)).astype(np.float32))      # pre-activations ~ N(0, 3)
    w2 = _bf16((rng.standard_normal((d, ff)) / math.sqrt(ff)).astype(np.float32))
    b1 = torch.from_numpy(np.linspace(-6, 6, ff).astype(np.float32))
    b2 = torch.zeros(d)
    ln = (torch.ones(d), torch.zeros(d))
    dv = _dev()
    xf = x.float().numpy()
    t = (xf @ w1.float().numpy().T + b1.numpy()).astype(np.float32)
    assert np.abs(t).max() > 10 and (np.abs(t) < 1e-2).any()
    h = _bf16(orc.gelu(t)).float().numpy()
    ref = orc.layer_norm((xf + h @ w2.float().numpy().T).astype(np.float32), ln[0].numpy(), ln[1].numpy())
    for ew in ("8", "16"):
        monkeypatch.setenv("SVOL_FFN_EPI_WARPS", ew)
        out = ops.ffn(x.to(dv), w1.to(dv), b1.to(dv), w2.to(dv), b2.to(dv), (ln[0].to(dv), ln[1].to(dv)))
        torch.cuda.synchronize()
        _assert_close(_f(out["out"]), ref, what=f"out ({ew} epilogue warps)")


@pytest.mark.parametrize("M,ff,posmode", FFN_CASES)
def test_ffn_fused(M, ff, posmode):
    """LayerNorm(x + fc2(GELU(fc1(x)))) -- cross_modal_transformer.py:142-143,163-179 -- vs the oracle's fp32 path
    on bf16-rounded operands; the hidden activation is rounded to bf16 on both sides."""
    from svol_b200 import ops
    d = 256
    rng = np.random.RandomState(M + ff)
    x = _bf16(rng.standard_normal((M, d)).astype(np.float32))
    w1 = _bf16((rng.standard_normal((ff, d)) / math.sqrt(d)).astype(np.float32))
    w2 = _bf16((rng.standard_normal((d, ff)) / math.sqrt(ff)).astype(np.float32))
    b1 = torch.from_numpy((0.5 * rng.standard_normal(ff)).astype(np.float32))
    b2 = torch.from_numpy((0.5 * rng.standard_normal(d)).astype(np.float32))
    ln = (torch.from_numpy((1 + 0.1 * rng.standard_normal(d)).astype(np.float32)),
          torch.from_numpy((0.1 * rng.standard_normal(d)).astype(np.float32)))
    pos_mod = 320 if posmode == "mod" else 0
    pos, theta = None, None
    dv = _dev()
    cu = lambda t: None if t is None else t.to(dv)
    if posmode == "theta":
        mask = np.ones((2, M // 2), np.float32)
        mask[1, -300:] = 0
        pos = torch.from_numpy(orc.position_embedding_sine(mask, d).reshape(M, d).astype(np.float32))
        theta = ops.posenc_theta(torch.from_numpy(mask).to(dv)).reshape(-1)
        out = ops.ffn(cu(x), cu(w1), cu(b1), cu(w2), cu(b2), (cu(ln[0]), cu(ln[1])), pos_theta=theta)
    else:
        if posmode is not None:
            pos = torch.from_numpy(rng.standard_normal((pos_mod or M, d)).astype(np.float32))
        out = ops.ffn(cu(x), cu(w1), cu(b1), cu(w2), cu(b2), (cu(ln[0]), cu(ln[1])), pos=cu(pos), pos_mod=pos_mod)
    torch.cuda.synchronize()
    xf = x.float().numpy()
    h = orc.gelu((xf @ w1.float().numpy().T + b1.numpy()).astype(np.float32))
    h = _bf16(h).float().numpy()                                   # the kernel keeps H in bf16
    y = xf + h @ w2.float().numpy().T + b2.numpy()
    ref = orc.layer_norm(y.astype(np.float32), ln[0].numpy(), ln[1].numpy())
    _assert_close(_f(out["out"]), ref, what="out")
    if pos is not None:
        prow = np.arange(M) % pos_mod if pos_mod else np.arange(M)
        _assert_close(_f(out["out_pos"]), ref + pos.numpy()[prow], what="out_pos")


@pytest.mark.parametrize("Lq,Lk,masked", [(1568, 1568, False), (320, 1568, True), (700, 320, False)])
def test_persistent_attention_kernel_matches_one_item_kernel(Lq, Lk, masked):
    """The persistent variant of the attention kernel (SVOL_ATTN_PERSISTENT, off by default: measured slower, see
    attn_tc.cu) walks several work items per CTA with cross-item barrier phases; its output must be bit-identical to the
    one-item-per-CTA kernel's (same tiles, same operation order), also when launched back to back."""
    import math
    from svol_b200 import ops
    B, H, d = 6, 8, 256
    DEV = torch.device("cuda:0")
    g = torch.Generator().manual_seed(3)
    q = (torch.randn(B * Lq, d, generator=g) * 2.0 * math.log2(math.e) / math.sqrt(32)).to(torch.bfloat16).to(DEV)
    k = torch.randn(B * Lk, d, generator=g).to(torch.bfloat16).to(DEV)
    pitch = (Lk + 7) // 8 * 8
    vt = torch.zeros(B * d, pitch, dtype=torch.bfloat16)
    vt[:, :Lk] = torch.randn(B * d, Lk, generator=g).to(torch.bfloat16)
    vt = vt.to(DEV)
    mask = None
    if masked:
        mask = torch.ones(B, Lk)
        for b in range(0, B, 2):
            mask[b, Lk - 49 * (1 + b):] = 0
        mask = mask.to(DEV)
    old = os.environ.get("SVOL_ATTN_PERSISTENT")
    old_small = os.environ.get("SVOL_ATTN_SMALL")
    try:
        os.environ["SVOL_ATTN_SMALL"] = "0"           # both sides through the tcgen05 kernels (the (700, 320) case is a short-key shape)
        os.environ["SVOL_ATTN_PERSISTENT"] = "0"
        ref = ops.attention(q, k, vt, B, H, Lq, Lk, key_mask=mask).clone()
        os.environ["SVOL_ATTN_PERSISTENT"] = "2"
        outs = [ops.attention(q, k, vt, B, H, Lq, Lk, key_mask=mask) for _ in range(8)]
        torch.cuda.synchronize()
    finally:
        if old is None:
            os.environ.pop("SVOL_ATTN_PERSISTENT", None)
        else:
            os.environ["SVOL_ATTN_PERSISTENT"] = old
        if old_small is None:
            os.environ.pop("SVOL_ATTN_SMALL", None)
        else:
            os.environ["SVOL_ATTN_SMALL"] = old_small
    # Rows of a last query tile with at most 64 rows go through the one-item kernel's duplicated-row walk (two row copies share
    # every half tile's keys: a different summation order), which the persistent variant does not have: those rows agree to
    # bf16 rounding, every other row bit for bit.
    tail = Lq % 256 if 0 < Lq % 256 <= 64 and Lk >= 4 * 128 else 0
    rows = torch.arange(B * Lq, device=DEV) % Lq
    body = rows < Lq - tail
    assert all(torch.equal(o, outs[0]) for o in outs)                      # repeatable, also back to back
    assert torch.equal(outs[0][body], ref[body])
    if tail:
        diff = (outs[0][~body].float() - ref[~body].float()).abs().max().item()
        assert diff <= 2e-2, diff


@pytest.mark.parametrize("Lq,Lk,masked", [(1568, 1568, False), (320, 1568, True), (320, 320, False), (290, 700, True)])
def test_looping_attention_ctas_match_one_cta_per_item(Lq, Lk, masked):
    """With more work items than SMs the inference attention kernel runs one looping CTA per SM (items handed out by a
    global counter, barriers re-initialised between items; SVOL_ATTN_LOOP=0 switches back to one CTA per item).  The
    items are processed by the same code, so the output must be bit-identical -- also over repeated launches (every launch
    leaves its item counter at zero) and under CUDA-graph replay (the captured launch keeps its counter slot)."""
    import math
    from svol_b200 import ops
    B, H, d = 16, 8, 256                      # 7 / 2 / 2 / 2 query-tile pairs x 8 heads x 16 samples: 896 / 256 items
    DEV = torch.device("cuda:0")
    g = torch.Generator().manual_seed(11)
    q = (torch.randn(B * Lq, d, generator=g) * 2.0 * math.log2(math.e) / math.sqrt(32)).to(torch.bfloat16).to(DEV)
    k = torch.randn(B * Lk, d, generator=g).to(torch.bfloat16).to(DEV)
    pitch = (Lk + 7) // 8 * 8
    vt = torch.zeros(B * d, pitch, dtype=torch.bfloat16)
    vt[:, :Lk] = torch.randn(B * d, Lk, generator=g).to(torch.bfloat16)
    vt = vt.to(DEV)
    mask = None
    if masked:
        mask = torch.ones(B, Lk)
        for b in range(0, B, 2):
            mask[b, Lk - 17 * (1 + b):] = 0
        mask = mask.to(DEV)
    old = os.environ.get("SVOL_ATTN_LOOP")
    old_small = os.environ.get("SVOL_ATTN_SMALL")
    try:
        os.environ["SVOL_ATTN_SMALL"] = "0"           # the (320, 320) case is a short-key shape: keep it on the tcgen05 kernel
        os.environ["SVOL_ATTN_LOOP"] = "0"
        ref = ops.attention(q, k, vt, B, H, Lq, Lk, key_mask=mask).clone()
        os.environ["SVOL_ATTN_LOOP"] = "1"
        outs = [ops.attention(q, k, vt, B, H, Lq, Lk, key_mask=mask).clone() for _ in range(6)]
        # graph replay of the looping kernel
        static_out = torch.empty_like(ref)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            ops.attention(q, k, vt, B, H, Lq, Lk, key_mask=mask, out=static_out)
        torch.cuda.current_stream().wait_stream(s)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            ops.attention(q, k, vt, B, H, Lq, Lk, key_mask=mask, out=static_out)
        for _ in range(3):
            static_out.zero_()
            graph.replay()
            outs.append(static_out.clone())
        torch.cuda.synchronize()
    finally:
        if old is None:
            os.environ.pop("SVOL_ATTN_LOOP", None)
        else:
            os.environ["SVOL_ATTN_LOOP"] = old
        if old_small is None:
            os.environ.pop("SVOL_ATTN_SMALL", None)
        else:
            os.environ["SVOL_ATTN_SMALL"] = old_small
    assert torch.isfinite(ref.float()).all()
    for i, o in enumerate(outs):
        assert torch.equal(o, ref), f"launch {i}: {(o.float() - ref.float()).abs().max().item()}"
