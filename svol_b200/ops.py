"""Functional wrappers over the C ABI (one Python function per ``svol_*`` entry point that takes
torch CUDA tensors).  The launch plan in ``engine.py`` records the same calls with pre-built argument
structures; these wrappers are the convenient form for tests, notebooks and one-off use."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import ACT_GELU, ACT_NONE, ACT_RELU  # noqa: F401  (re-exported)

_P = _lib.ptr


def gemm(A: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor] = None, act: int = ACT_NONE,
         residual: Optional[torch.Tensor] = None, ln: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
         pos: Optional[torch.Tensor] = None, pos_mod: int = 0, want_out: bool = True, vt_len: int = 0,
         plain: bool = False, eps: float = 1e-5, pos_theta: Optional[torch.Tensor] = None,
         A2: Optional[torch.Tensor] = None, split_block: int = 0):
    """out = epilogue(A @ W.T).  A [M,K] bf16, W [N,K] bf16.  Returns a dict with 'out' [M,N] bf16,
    'out_pos' (when ``pos`` is given) and 'out_vt' [(M/vt_len)*N, round_up(vt_len,8)] (when ``vt_len``)."""
    _lib.require_device()
    lib = _lib.get_lib()
    M, K = A.shape
    N = W.shape[0]
    res = {}
    a = _lib.GemmArgs()
    a.A, a.W, a.M, a.N, a.K, a.lda, a.ldw = _P(A), _P(W), M, N, K, A.stride(0), W.stride(0)
    e = a.ep
    e.bias, e.act = _P(bias), act
    if residual is not None:
        e.residual, e.ld_res = _P(residual), residual.stride(0)
    if ln is not None:
        e.ln_weight, e.ln_bias, e.ln_eps = _P(ln[0]), _P(ln[1]), eps
    n_out = split_block * 256 if split_block else N       # split launch: out covers the blocks before the split
    if A2 is not None:
        a.A2, a.lda2, a.split_block = _P(A2), A2.stride(0), split_block
    e.ld_out = n_out
    if want_out:
        res["out"] = torch.empty((M, n_out), device=A.device, dtype=torch.bfloat16)
        e.out = _P(res["out"])
    if pos is not None:
        res["out_pos"] = torch.empty((M, N), device=A.device, dtype=torch.bfloat16)
        e.out_pos, e.pos, e.ld_pos, e.pos_row_mod = _P(res["out_pos"]), _P(pos), pos.stride(0), pos_mod
    elif pos_theta is not None:
        res["out_pos"] = torch.empty((M, N), device=A.device, dtype=torch.bfloat16)
        e.out_pos, e.pos_theta = _P(res["out_pos"]), _P(pos_theta)
    if vt_len:
        pitch = (vt_len + 7) // 8 * 8
        res["out_vt"] = torch.zeros(((M // vt_len) * (N - (n_out if split_block else 0)), pitch), device=A.device,
                                    dtype=torch.bfloat16)
        e.out_vt, e.vt_len, e.vt_pitch = _P(res["out_vt"]), vt_len, pitch
    fn = lib.svol_gemm_bf16_plain if plain else lib.svol_gemm_bf16
    _lib.check(fn(C.byref(a), _lib.stream_ptr()), "gemm")
    return res


def ffn(x: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor,
        ln: Tuple[torch.Tensor, torch.Tensor], pos: Optional[torch.Tensor] = None, pos_mod: int = 0, eps: float = 1e-5,
        pos_theta: Optional[torch.Tensor] = None):
    """LayerNorm(x + fc2(GELU(fc1(x)))) for x [M,256] bf16, w1 [ff,256] bf16, w2 [256,ff] bf16.
    Returns {'out': [M,256] bf16, 'out_pos': out + pos (when ``pos`` is given)}."""
    _lib.require_device()
    M, d = x.shape
    res = {"out": torch.empty((M, d), device=x.device, dtype=torch.bfloat16)}
    a = _lib.FfnArgs()
    a.x, a.w1, a.b1, a.w2, a.b2 = _P(x), _P(w1), _P(b1), _P(w2), _P(b2)
    a.ln_weight, a.ln_bias, a.out = _P(ln[0]), _P(ln[1]), _P(res["out"])
    a.M, a.d, a.ff, a.ldx, a.ldw1, a.ldw2, a.ld_out = M, d, w1.shape[0], x.stride(0), w1.stride(0), w2.stride(0), d
    a.ln_eps = eps
    if pos is not None:
        res["out_pos"] = torch.empty((M, d), device=x.device, dtype=torch.bfloat16)
        a.out_pos, a.pos, a.ld_pos, a.pos_row_mod = _P(res["out_pos"]), _P(pos), pos.stride(0), pos_mod
    elif pos_theta is not None:
        res["out_pos"] = torch.empty((M, d), device=x.device, dtype=torch.bfloat16)
        a.out_pos, a.pos_theta = _P(res["out_pos"]), _P(pos_theta)
    _lib.check(_lib.get_lib().svol_ffn_bf16(C.byref(a), _lib.stream_ptr()), "ffn")
    return res


def attention(q: torch.Tensor, k: torch.Tensor, vt: torch.Tensor, B: int, H: int, Lq: int, Lk: int,
              key_mask: Optional[torch.Tensor] = None, plain: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q [B*Lq, >=H*32] (pre-scaled by log2(e)/sqrt(32)), k [B*Lk, >=H*32], vt [B*H*32, pitch] bf16.  ``out``: an existing
    [B*Lq, H*32] bf16 tensor to write into (stable address under CUDA-graph capture)."""
    _lib.require_device()
    lib = _lib.get_lib()
    if out is None:
        out = torch.empty((B * Lq, H * 32), device=q.device, dtype=torch.bfloat16)
    a = _lib.AttnArgs()
    a.q, a.k, a.vt, a.key_mask, a.out = _P(q), _P(k), _P(vt), _P(key_mask), _P(out)
    a.B, a.H, a.Lq, a.Lk = B, H, Lq, Lk
    a.ldq, a.ldk, a.ldo, a.vt_pitch = q.stride(0), k.stride(0), out.stride(0), vt.stride(0)
    fn = lib.svol_attention_bf16_plain if plain else lib.svol_attention_bf16
    _lib.check(fn(C.byref(a), _lib.stream_ptr()), "attention")
    return out


def posenc_theta(mask: torch.Tensor) -> torch.Tensor:
    """mask [B,L] float -> theta [B,L] fp32 (angles of the normalised sine positional encoding)."""
    _lib.require_device()
    B, L = mask.shape
    theta = torch.empty((B, L), device=mask.device, dtype=torch.float32)
    _lib.check(_lib.get_lib().svol_posenc_theta(_P(mask), _P(theta), B, L, _lib.stream_ptr()), "posenc_theta")
    return theta


def layernorm_to_bf16(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float = 1e-5, drop_p: float = 0.0, seed=None,
                      site: int = 0) -> torch.Tensor:
    _lib.require_device()
    rows, cols = x.shape
    y = torch.empty((rows, cols), device=x.device, dtype=torch.bfloat16)
    if drop_p > 0:
        _lib.check(_lib.get_lib().svol_layernorm_f32_to_bf16_dropout(_P(x), _P(w), _P(b), _P(y), rows, cols, eps, drop_p, _P(seed), site,
                                                                     _lib.stream_ptr()), "layernorm_dropout")
    else:
        _lib.check(_lib.get_lib().svol_layernorm_f32_to_bf16(_P(x), _P(w), _P(b), _P(y), rows, cols, eps, _lib.stream_ptr()),
                   "layernorm")
    return y


def ln_linear_f32(x, ln_w, ln_b, w, b, relu: bool, eps: float = 1e-5, drop_p: float = 0.0, seed=None, site: int = 0) -> torch.Tensor:
    _lib.require_device()
    rows, in_dim = x.shape
    out_dim = w.shape[0]
    y = torch.empty((rows, out_dim), device=x.device, dtype=torch.float32)
    if drop_p > 0:
        _lib.check(_lib.get_lib().svol_ln_linear_f32_dropout(_P(x), _P(ln_w), _P(ln_b), _P(w), _P(b), int(relu), _P(y), rows, in_dim,
                                                             out_dim, eps, drop_p, _P(seed), site, _lib.stream_ptr()), "ln_linear_dropout")
    else:
        _lib.check(_lib.get_lib().svol_ln_linear_f32(_P(x), _P(ln_w), _P(ln_b), _P(w), _P(b), int(relu), _P(y), rows, in_dim,
                                                     out_dim, eps, _lib.stream_ptr()), "ln_linear")
    return y


def posenc_sine(mask: torch.Tensor, d: int) -> torch.Tensor:
    _lib.require_device()
    B, L = mask.shape
    pos = torch.empty((B, L, d), device=mask.device, dtype=torch.float32)
    _lib.check(_lib.get_lib().svol_posenc_sine(_P(mask), _P(pos), B, L, d, _lib.stream_ptr()), "posenc")
    return pos


def add_pos_bf16(x: torch.Tensor, pos: Optional[torch.Tensor], rows: int, mod: int = 0) -> torch.Tensor:
    _lib.require_device()
    cols = x.shape[-1]
    out = torch.empty((rows, cols), device=x.device, dtype=torch.bfloat16)
    _lib.check(_lib.get_lib().svol_add_pos_bf16(_P(x), _P(pos), _P(out), rows, cols, mod, _lib.stream_ptr()), "add_pos")
    return out


def gate(x: torch.Tensor, xpos: torch.Tensor, sketch: torch.Tensor, in_w: torch.Tensor, in_b: torch.Tensor,
         ln_w: torch.Tensor, ln_b: torch.Tensor, pos: torch.Tensor, B: int, L: int, H: int = 8, eps: float = 1e-5,
         pos_is_theta: bool = False):
    """The three gate kernels in sequence.  Returns (mem, mem_pos, att [B,L], scores [B,H,L]).  ``pos`` is the fp32
    table [B*L, d], or (``pos_is_theta``) the [B*L] angles of :func:`posenc_theta`."""
    _lib.require_device()
    lib = _lib.get_lib()
    d = x.shape[-1]
    s = _lib.stream_ptr()
    u = torch.empty((B, H, d), device=x.device, dtype=torch.float32)
    scores = torch.empty((B, H, L), device=x.device, dtype=torch.float32)
    att = torch.empty((B, L), device=x.device, dtype=torch.float32)
    mem, mem_pos = torch.empty_like(x), torch.empty_like(x)
    _lib.check(lib.svol_gate_vectors(_P(sketch), _P(in_w), _P(in_b), _P(u), B, d, H, s), "gate_vectors")
    _lib.check(lib.svol_gate_scores(_P(xpos), _P(u), _P(scores), B, L, d, H, s), "gate_scores")
    fn = lib.svol_gate_apply_theta if pos_is_theta else lib.svol_gate_apply
    _lib.check(fn(_P(x), _P(scores), _P(ln_w), _P(ln_b), _P(pos), _P(mem), _P(mem_pos), _P(att), B, L, d, H, eps, s),
               "gate_apply")
    return mem, mem_pos, att, scores


def gate_fused(x: torch.Tensor, sketch: torch.Tensor, in_w: torch.Tensor, in_b: torch.Tensor, ln_w: torch.Tensor,
               ln_b: torch.Tensor, theta: torch.Tensor, B: int, L: int, H: int = 8, eps: float = 1e-5):
    """svol_gate_vectors + the one-launch gate (svol_gate_fused).  Returns (mem, mem_pos, att [B,L], scores [B,H,L])."""
    _lib.require_device()
    lib = _lib.get_lib()
    d = x.shape[-1]
    s = _lib.stream_ptr()
    if not lib.svol_gate_fused_supported(L):
        raise ValueError(f"svol_gate_fused: L = {L} tokens per sample do not fit one cluster's shared memory")
    u = torch.empty((B, H, d), device=x.device, dtype=torch.float32)
    scores = torch.empty((B, H, L), device=x.device, dtype=torch.float32)
    att = torch.empty((B, L), device=x.device, dtype=torch.float32)
    mem, mem_pos = torch.empty_like(x), torch.empty_like(x)
    _lib.check(lib.svol_gate_vectors(_P(sketch), _P(in_w), _P(in_b), _P(u), B, d, H, s), "gate_vectors")
    _lib.check(lib.svol_gate_fused(_P(x), _P(u), _P(ln_w), _P(ln_b), _P(theta), _P(mem), _P(mem_pos), _P(att), _P(scores),
                                   B, L, d, H, eps, s), "gate_fused")
    return mem, mem_pos, att, scores


def heads(hs: torch.Tensor, h2: torch.Tensor, wc, bc, wb, bb):
    _lib.require_device()
    rows, d = hs.shape
    logits = torch.empty((rows, 2), device=hs.device, dtype=torch.float32)
    boxes = torch.empty((rows, 4), device=hs.device, dtype=torch.float32)
    _lib.check(_lib.get_lib().svol_heads(_P(hs), _P(h2), _P(wc), _P(bc), _P(wb), _P(bb), _P(logits), _P(boxes), rows, d,
                                         _lib.stream_ptr()), "heads")
    return logits, boxes


# ---------------------------------------------------------------------------------------------------------------
# training-step entry points (functional form, for tests and one-off use; the training plan is train_engine.py)
# ---------------------------------------------------------------------------------------------------------------
def attention_train(q, k, vt, B, H, Lq, Lk, key_mask=None):
    """Like :func:`attention`, also returns the base-2 log-sum-exp [B, H, round_up(Lq, 64)] (+inf padded)."""
    _lib.require_device()
    out = torch.empty((B * Lq, H * 32), device=q.device, dtype=torch.bfloat16)
    pitch = (Lq + 63) // 64 * 64
    lse = torch.full((B, H, pitch), float("inf"), device=q.device, dtype=torch.float32)
    a = _lib.AttnArgs()
    a.q, a.k, a.vt, a.key_mask, a.out = _P(q), _P(k), _P(vt), _P(key_mask), _P(out)
    a.B, a.H, a.Lq, a.Lk = B, H, Lq, Lk
    a.ldq, a.ldk, a.ldo, a.vt_pitch = q.stride(0), k.stride(0), out.stride(0), vt.stride(0)
    a.lse, a.lse_pitch = _P(lse), pitch
    _lib.check(_lib.get_lib().svol_attention_bf16(C.byref(a), _lib.stream_ptr()), "attention")
    return out, lse


def attention_backward(q, k, v, kt, qt, o, d_o, d_ot, lse, B, H, Lq, Lk, key_mask=None):
    """Returns (dq [B*Lq,256] w.r.t. the UNSCALED query projection, dk, dv [B*Lk,256]) bf16."""
    _lib.require_device()
    dq = torch.empty((B * Lq, H * 32), device=q.device, dtype=torch.bfloat16)
    dk = torch.empty((B * Lk, H * 32), device=q.device, dtype=torch.bfloat16)
    dv = torch.empty_like(dk)
    delta = torch.zeros_like(lse)
    a = _lib.AttnBwdArgs()
    a.q, a.k, a.v, a.kt, a.qt, a.o, a.d_o, a.d_ot = _P(q), _P(k), _P(v), _P(kt), _P(qt), _P(o), _P(d_o), _P(d_ot)
    a.lse, a.delta, a.key_mask, a.dq, a.dk, a.dv = _P(lse), _P(delta), _P(key_mask), _P(dq), _P(dk), _P(dv)
    a.B, a.H, a.Lq, a.Lk = B, H, Lq, Lk
    a.ldq, a.ldk, a.ldv, a.ld_o, a.ld_do = q.stride(0), k.stride(0), v.stride(0), o.stride(0), d_o.stride(0)
    a.ld_dq, a.ld_dk, a.ld_dv = dq.stride(0), dk.stride(0), dv.stride(0)
    a.kt_pitch, a.qt_pitch, a.stat_pitch = kt.stride(0), qt.stride(0), lse.shape[-1]
    _lib.check(_lib.get_lib().svol_attention_backward_bf16(C.byref(a), _lib.stream_ptr()), "attention_backward")
    return dq, dk, dv


def layernorm_bf16(z, w, b, pos=None, pos_mod=0, theta=None, eps: float = 1e-5, drop_p: float = 0.0, seed=None, site: int = 0):
    """``seed``: int64 device tensor of one element (dropout, see include/svol_b200.h)."""
    _lib.require_device()
    y = torch.empty_like(z)
    y_pos = torch.empty_like(z) if (pos is not None or theta is not None) else None
    _lib.check(_lib.get_lib().svol_layernorm_bf16(_P(z), _P(w), _P(b), _P(y), _P(y_pos), _P(pos), pos_mod, _P(theta), z.shape[0],
                                                  z.shape[1], eps, drop_p, _P(seed), site, _lib.stream_ptr()), "layernorm_bf16")
    return y, y_pos


def layernorm_backward(z, dys, gamma, att=None, want_dx: bool = True, eps: float = 1e-5, drop_p: float = 0.0, seed=None,
                       site: int = 0):
    """Returns (dx bf16 | None, datt | None, dgamma, dbeta)."""
    _lib.require_device()
    rows, cols = z.shape
    dys = list(dys) + [None] * (3 - len(dys))
    dx = torch.empty((rows, cols), device=z.device, dtype=torch.bfloat16) if want_dx else None
    datt = torch.empty(rows, device=z.device, dtype=torch.float32) if att is not None else None
    dg, db = torch.zeros(cols, device=z.device), torch.zeros(cols, device=z.device)
    _lib.check(_lib.get_lib().svol_layernorm_backward(_P(z), int(z.dtype == torch.float32), _P(att), _P(dys[0]), _P(dys[1]), _P(dys[2]),
                                                      _P(gamma), _P(dx), _P(datt), _P(dg), _P(db), rows, cols, eps, drop_p,
                                                      _P(seed), site, _lib.stream_ptr()), "layernorm_backward")
    return dx, datt, dg, db


def gelu_bf16(x):
    _lib.require_device()
    y = torch.empty_like(x)
    _lib.check(_lib.get_lib().svol_gelu_bf16(_P(x), _P(y), x.numel(), _lib.stream_ptr()), "gelu")
    return y


def act_backward(dy, saved, mode):
    _lib.require_device()
    out = torch.empty_like(dy)
    _lib.check(_lib.get_lib().svol_act_backward(_P(dy), _P(saved), _P(out), dy.numel(), mode, _lib.stream_ptr()), "act_backward")
    return out


def transpose_bf16(x, want_colsum: bool = False):
    """x [rows, cols] bf16 -> (x^T [cols, round_up(rows, 64)] zero padded, colsum fp32 | None)."""
    _lib.require_device()
    rows, cols = x.shape
    rp = (rows + 63) // 64 * 64
    out = torch.zeros((cols, rp), device=x.device, dtype=torch.bfloat16)
    cs = torch.zeros(cols, device=x.device, dtype=torch.float32) if want_colsum else None
    _lib.check(_lib.get_lib().svol_transpose_bf16(_P(x), x.stride(0), rows, cols, _P(out), rp, _P(cs), _lib.stream_ptr()), "transpose")
    return out, cs
