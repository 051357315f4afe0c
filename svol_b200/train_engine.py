"""Launch plans of the SVANet head TRAINING step on the CUDA C ABI: a forward that keeps what the backward
needs, and the explicit backward the reference gets from ``loss.backward()`` (train.py:216-232).

Differences from the inference plan (``engine.py``):

  * every LayerNorm input ``z`` (the residual sum in front of norm1..norm6) is stored, so the projections run
    without the fused LayerNorm epilogue and ``svol_layernorm_bf16`` follows them;
  * the FFN runs as two GEMMs (fc1 stores its pre-activation AND GELU(pre) from one epilogue, then fc2): the fused
    kernel of ``ffn_tc.cu`` keeps the 2048-wide hidden activation on the SM, the backward needs it (``gelu'``, ``dW2``);
  * q, k, v projections are three launches that also write the per-head transposed copies (Q^T, K^T, V^T) the
    attention backward's MMAs take as K-major operands; the attention forward stores each row's log-sum-exp.

Backward, per ``nn.Linear``: ``dX = dY W`` is the tcgen05 GEMM with the transposed weight as its W operand,
``dW = dY^T X`` the same GEMM on the untransposed operands (MN-major shared-memory descriptors; the contraction over
the token rows is split over the SMs and accumulated into the fp32 gradient with vector atomics), ``db`` a column sum.  LayerNorm / GELU / ReLU / gate / heads backward are the kernels of ``train.cu``,
the attention backward is ``attn_bwd_tc.cu``.  Activation gradients are bf16, parameter gradients fp32 views of one
flat buffer (``grad_of(param)``), zeroed at the start of every backward.

Dropout: the reference applies ``Dropout(input_dropout)`` after the LayerNorm of each input-projection stage in
train mode (svanet.py:168-170: four sites, frame tokens and sketch, two stages each).  The forward kernels apply a
counter-based mask (splitmix64 of element index, step seed and site; see include/svol_b200.h), the backward kernels
recompute it from the same seed -- nothing is stored.  The step seed lives in a device scalar that the host bumps
before every training forward, so the plans still replay as CUDA graphs.  The random stream is this library's own
(torch's Philox stream cannot be reproduced); parity tests rebuild the masks on the host from the same hash.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, List

import torch

from . import _lib
from ._lib import ACT_GELU, ACT_NONE, ACT_RELU, AttnArgs, AttnBwdArgs, GemmArgs
from .engine import HEAD_DIM, LN_EPS, WeightPacker, _Plan, _round_up

P = _lib.ptr


class TrainEngine:
    """Forward-for-training + backward of one ``SVANet`` module.  Shares the packed weights of the module's
    ``HeadEngine`` and adds the transposed / unscaled copies the backward GEMMs need."""

    def __init__(self, module):
        self.module = module
        self.head = module.engine
        self._wt: Dict[str, torch.Tensor] = {}
        self._wt_state = None
        self._packer = WeightPacker()
        # both plans replay as CUDA graphs after one eager warm-up run per shape (SVOL_B200_TRAIN_GRAPH=0: eager launches)
        self.use_graph = os.environ.get("SVOL_B200_TRAIN_GRAPH", "1") != "0"
        # True: loss.backward() fills param.grad like the reference (one copy per parameter and step).  A loop that
        # feeds FusedAdamW.step(from_engine=True) from grad_flat can switch it off.
        self.publish_grads = True
        self.seed_base = int(torch.initial_seed()) & 0x3FFFFFFF      # dropout stream; step k uses seed_base + k
        self.step_index = 0
        self.last_seed = None
        self._side = None
        self.forward_token = 0
        self._plans: Dict[tuple, dict] = {}
        self._params: List[torch.nn.Parameter] = []
        self.grad_flat: torch.Tensor = None
        self._grad_views: Dict[int, torch.Tensor] = {}

    # ------------------------------------------------------------------ parameters / gradients
    def _setup_grads(self):
        params = list(self.module.parameters())
        if self.grad_flat is not None and len(params) == len(self._params) and all(a is b for a, b in zip(params, self._params)) \
                and self.grad_flat.device == params[0].device:
            return
        dev = params[0].device
        total = sum(_round_up(p.numel(), 4) for p in params)
        self.grad_flat = torch.zeros(total, device=dev, dtype=torch.float32)
        self._params, self._grad_views, off = params, {}, 0
        for p in params:
            self._grad_views[id(p)] = self.grad_flat[off:off + p.numel()].view_as(p)
            off += _round_up(p.numel(), 4)
        self._plans.clear()

    def grad_of(self, p: torch.nn.Parameter) -> torch.Tensor:
        return self._grad_views[id(p)]

    @torch.no_grad()
    def _pack_transposed(self):
        """bf16 W^T operands of the dgrad GEMMs (unscaled q rows), refreshed with the forward's packed weights."""
        w = self.head._weights()
        # An explicit generation counter, not the (data_ptr, _version) tuple: FusedAdamW updates the flat parameter buffer
        # through a raw pointer, which changes neither, and only resets the head engine's state -- the forward weights
        # were then repacked while these transposed copies silently kept the initial values (round-1 ADVICE, high).
        st = self.head.weights_generation
        if st == self._wt_state and self._wt:
            return w
        m = self.module
        d = m.transformer.d_model
        bf = torch.bfloat16
        pk = self._packer
        pk.begin(next(m.parameters()).device)

        def put(name, param, row0=0, rows=None):
            pk.put(name, param, bf, row0=row0, rows=rows, transpose=True)

        for i, lin in enumerate(m.input_video_proj):
            put(f"in_video.{i}.wT", lin.net[1].weight)
        for li, layer in enumerate(m.transformer.layers):
            p = f"l{li}."
            for tag, att in (("sa", layer.content_self_attn), ("ta", layer.token_self_attn), ("ca", layer.content_token_cross_attn)):
                W = att.in_proj_weight
                put(p + tag + ".wqkT", W, rows=2 * d)      # [d, 2d]: d(x+pos) = [dq | dk] [Wq ; Wk]
                put(p + tag + ".wqT", W, rows=d)
                put(p + tag + ".wkT", W, row0=d, rows=d)
                put(p + tag + ".wvT", W, row0=2 * d)
                put(p + tag + ".woT", att.out_proj.weight)
            for tag, mlp in (("mlp1", layer.mlp1), ("mlp2", layer.mlp2)):
                put(p + tag + ".w1T", mlp.fc1.weight)      # [d, ff]
                put(p + tag + ".w2T", mlp.fc2.weight)      # [ff, d]
        for i in range(2):
            put(f"box.{i}.wT", m.bbox_embed.layers[i].weight)
        if pk.end():
            self._plans.clear()
        self._wt = pk.tensors
        self._wt_state = st
        return w

    # ------------------------------------------------------------------ plans
    def _build(self, B: int, L: int, d_in: int, want_dx: bool = False) -> dict:
        m, w, wt = self.module, self.head._w, self._wt
        lib = _lib.get_lib()
        dev = w["cls.w"].device
        d, H, ff = 256, 8, m.transformer.layers[0].mlp1.fc1.weight.shape[0]
        Q, NL = m.num_queries, len(m.transformer.layers)
        n_proj = len(m.input_video_proj)
        if n_proj != 2:
            raise NotImplementedError("training plan is written for n_input_proj == 2 (the shipped configuration)")
        d_sk = m.input_sketch_proj[0].net[1].weight.shape[1]
        M, MQ = B * L, B * Q
        Lp, Qp = _round_up(L, 8), _round_up(Q, 8)
        Ls, Qs = _round_up(L, 64), _round_up(Q, 64)          # pitch of the per-row softmax statistics
        bf, f32 = torch.bfloat16, torch.float32
        fwd, bwd = _Plan(), _Plan()
        bufs: Dict[str, torch.Tensor] = {}
        G = self.grad_of

        def buf(name, shape, dtype, fill=None):
            if fill is None:
                t = torch.empty(shape, device=dev, dtype=dtype)
            else:
                t = torch.full(shape, fill, device=dev, dtype=dtype)
            bufs[name] = t
            return t

        # ---------------------------------------------------------------- call recorders
        def gemm(plan, name, A, W, bias=None, out=None, act=ACT_NONE, residual=None, out_pos=None, pos_t=None, pos_mod=0,
                 theta_t=None, out_vt=None, vt_len=0, vt_pitch=0, out_f32=None, out_pre=None, dact=None, dact_mode=0):
            a = GemmArgs()
            a.A, a.W = P(A), P(W)
            a.M, a.K = A.shape
            a.N = W.shape[0]
            assert W.shape[1] == a.K, (name, A.shape, W.shape)
            a.lda, a.ldw = A.stride(0), W.stride(0)
            e = a.ep
            e.bias, e.act = P(bias), act
            e.residual, e.ld_res = P(residual), (residual.stride(0) if residual is not None else 0)
            e.out, e.out_pos = P(out), P(out_pos)
            e.ld_out = out.stride(0) if out is not None else (out_pos.stride(0) if out_pos is not None else 0)
            if out_pos is not None and theta_t is not None:
                e.pos_theta = P(theta_t)
            elif out_pos is not None:
                e.pos, e.ld_pos, e.pos_row_mod = P(pos_t), pos_t.stride(0), pos_mod
            if out_vt is not None:
                e.out_vt, e.vt_len, e.vt_pitch = P(out_vt), vt_len, vt_pitch
            if out_pre is not None:                      # fc1 keeps its pre-activation next to GELU(pre)
                e.out_pre = P(out_pre)
            if dact is not None:                         # activation backward fused into the dgrad epilogue
                e.dact_src, e.ld_dact, e.dact_mode = P(dact), dact.stride(0), dact_mode
            if out_f32 is not None:                      # weight-gradient mode: fp32 accumulation, split contraction
                assert out_f32.is_contiguous() and out_f32.shape == (a.M, a.N), name
                a.out_f32, a.ld_f32 = P(out_f32), out_f32.stride(0)
            plan.keep.append(a)
            plan.calls.append((name, lib.svol_gemm_bf16, (C.byref(a),)))

        def call(plan, name, fn, *args):
            plan.calls.append((name, fn, args))

        drop_p = float(getattr(m, "input_dropout", 0.0))
        seed = buf("seed", (1,), torch.int64, fill=0)

        def ln(name, z, g, b_, y, y_pos=None, pos_t=None, mod=0, theta_t=None, site=None):
            call(fwd, name, lib.svol_layernorm_bf16, P(z), P(g), P(b_), P(y), P(y_pos), P(pos_t), mod, P(theta_t), z.shape[0], d,
                 LN_EPS, drop_p if site is not None else 0.0, P(seed), site or 0)

        def ln_bwd(name, z, dy1, gamma, norm, dx, dy2=None, dy3=None, att=None, datt=None, z_f32=False, cols=256, site=None):
            call(bwd, name, lib.svol_layernorm_backward, P(z), int(z_f32), P(att), P(dy1), P(dy2), P(dy3), P(gamma),
                 P(dx), P(datt), P(G(norm.weight)), P(G(norm.bias)), dy1.shape[0], cols, LN_EPS,
                 drop_p if site is not None else 0.0, P(seed), site or 0)

        def attention(name, q, k, vt_t, out, lse, Lq, Lk, pitch, stat_pitch, mask=None):
            a = AttnArgs()
            a.q, a.k, a.vt, a.key_mask, a.out = P(q), P(k), P(vt_t), P(mask), P(out)
            a.B, a.H, a.Lq, a.Lk = B, H, Lq, Lk
            a.ldq, a.ldk, a.ldo, a.vt_pitch = q.stride(0), k.stride(0), out.stride(0), pitch
            a.lse, a.lse_pitch = P(lse), stat_pitch
            fwd.keep.append(a)
            fwd.calls.append((name, lib.svol_attention_bf16, (C.byref(a),)))

        def attention_bwd(name, q, k, v, kt, qt, o, d_o, d_ot, lse, delta, dq, dk, dv, Lq, Lk, kt_pitch, qt_pitch, stat_pitch,
                          mask=None):
            a = AttnBwdArgs()
            a.q, a.k, a.v, a.kt, a.qt, a.o, a.d_o, a.d_ot = P(q), P(k), P(v), P(kt), P(qt), P(o), P(d_o), P(d_ot)
            a.lse, a.delta, a.key_mask, a.dq, a.dk, a.dv = P(lse), P(delta), P(mask), P(dq), P(dk), P(dv)
            a.B, a.H, a.Lq, a.Lk = B, H, Lq, Lk
            a.ldq, a.ldk, a.ldv, a.ld_o, a.ld_do = q.stride(0), k.stride(0), v.stride(0), o.stride(0), d_o.stride(0)
            a.ld_dq, a.ld_dk, a.ld_dv = dq.stride(0), dk.stride(0), dv.stride(0)
            a.kt_pitch, a.qt_pitch, a.stat_pitch = kt_pitch, qt_pitch, stat_pitch
            bwd.keep.append(a)
            bwd.calls.append((name, lib.svol_attention_backward_bf16, (C.byref(a),)))

        R = NL * MQ

        def linear_bwd(name, dY, X, weight_grad, bias_grad, wT=None, dX=None, residual=None, out_vt=None, vt_len=0, vt_pitch=0,
                       dact=None, dact_mode=0):
            """dY [rows, N_out], X [rows, K_in] (bf16).  weight_grad [N_out, K_in] / bias_grad [N_out] fp32 views (accumulated).
            dW += dY^T X runs on the UNtransposed operands (MN-major descriptors, contraction split over the SMs);
            optional dgrad dX = dY W (+ residual, * f'(dact)), optionally also stored per-head transposed."""
            rows, n_out = dY.shape
            k_in = X.shape[1]
            assert X.shape[0] == rows and weight_grad.is_contiguous() and tuple(weight_grad.shape) == (n_out, k_in), name
            call(bwd, name + ".db", lib.svol_colsum_bf16, P(dY), dY.stride(0), rows, n_out, P(bias_grad))
            a = GemmArgs()
            a.A, a.W, a.M, a.N, a.K, a.lda, a.ldw = P(dY), P(X), n_out, k_in, rows, dY.stride(0), X.stride(0)
            a.out_f32, a.ld_f32, a.mn_major = P(weight_grad), k_in, 1
            bwd.keep.append(a)
            bwd.calls.append((name + ".wgrad", lib.svol_gemm_bf16, (C.byref(a),)))
            if dX is not None or out_vt is not None:
                gemm(bwd, name + ".dgrad", dY, wT, out=dX, residual=residual, out_vt=out_vt, vt_len=vt_len, vt_pitch=vt_pitch,
                     dact=dact, dact_mode=dact_mode)

        def bcall(name, fn, *args):
            call(bwd, name, fn, *args)

        def bstruct(name, fn, struct):
            bwd.keep.append(struct)
            bwd.calls.append((name, fn, (C.byref(struct),)))

        # ================================================================= FORWARD (training)
        x_in = buf("src_video", (B, L, d_in), f32)
        s_in = buf("src_sketch", (B, d_sk), f32)
        vmask = buf("src_video_mask", (B, L), f32)
        theta = buf("theta", (M,), f32)
        xn0 = buf("xn0", (M, d_in), bf)
        a0 = buf("a0", (M, d), bf)
        xn1 = buf("xn1", (M, d), bf)
        layers_in_X = [buf("X0", (M, d), bf)]
        layers_in_Xp = [buf("Xp0", (M, d), bf)]
        vp, sp = m.input_video_proj, m.input_sketch_proj
        fwd.ln_in_index = 0
        # dropout sites: 0 / 1 = frame-token stages, 2 / 3 = sketch stages (svanet.py:49-60,168-170)
        call(fwd, "ln_in", lib.svol_layernorm_f32_to_bf16_dropout, P(x_in), P(w["in_video.0.ln_w"]), P(w["in_video.0.ln_b"]), P(xn0),
             M, d_in, LN_EPS, drop_p, P(seed), 0)
        call(fwd, "posenc_theta", lib.svol_posenc_theta, P(vmask), P(theta), B, L)
        gemm(fwd, "in_proj0", xn0, w["in_video.0.w"], w["in_video.0.b"], out=a0, act=ACT_RELU)
        ln("in_ln1", a0, w["in_video.1.ln_w"], w["in_video.1.ln_b"], xn1, site=1)
        gemm(fwd, "in_proj1", xn1, w["in_video.1.w"], w["in_video.1.b"], out=layers_in_X[0], out_pos=layers_in_Xp[0], theta_t=theta)
        sk0, sk1 = buf("sk0", (B, d), f32), buf("sk1", (B, d), f32)
        call(fwd, "sk_proj0", lib.svol_ln_linear_f32_dropout, P(s_in), P(w["in_sketch.0.ln_w"]), P(w["in_sketch.0.ln_b"]),
             P(w["in_sketch.0.w"]), P(w["in_sketch.0.b"]), 1, P(sk0), B, d_sk, d, LN_EPS, drop_p, P(seed), 2)
        call(fwd, "sk_proj1", lib.svol_ln_linear_f32_dropout, P(sk0), P(w["in_sketch.1.ln_w"]), P(w["in_sketch.1.ln_b"]),
             P(w["in_sketch.1.w"]), P(w["in_sketch.1.b"]), 0, P(sk1), B, d, d, LN_EPS, drop_p, P(seed), 3)
        qe_bf = buf("qe_bf", (MQ, d), bf)
        zeros_q = buf("zeros_q", (MQ, d), bf, fill=0)
        call(fwd, "qe_bcast", lib.svol_add_pos_bf16, P(w["query_embed"]), None, P(qe_bf), MQ, d, Q)
        fwd.branch["qe_bcast"] = "q"
        hs = buf("hs", (NL, MQ, d), bf)
        S: List[dict] = []
        out_cur, outp_cur = zeros_q, qe_bf
        for li, layer in enumerate(m.transformer.layers):
            p = f"l{li}."
            s: Dict[str, torch.Tensor] = {}

            def sb(name, shape, dtype=bf, fill=None, _p=p, _s=s):
                t = buf(_p + name, shape, dtype, fill)
                _s[name] = t
                return t

            X, Xp = layers_in_X[li], layers_in_Xp[li]
            sb("u", (B, H, d), f32); sb("scores", (B, H, L), f32); sb("att", (B, L), f32)
            sb("mem", (M, d)); sb("memp", (M, d))
            call(fwd, p + "gate_vec", lib.svol_gate_vectors, P(sk1), P(w[p + "gate.w"]), P(w[p + "gate.b"]), P(s["u"]), B, d, H)
            call(fwd, p + "gate_scores", lib.svol_gate_scores, P(Xp), P(s["u"]), P(s["scores"]), B, L, d, H)
            call(fwd, p + "gate_apply", lib.svol_gate_apply_theta, P(X), P(s["scores"]), P(w[p + "n1.w"]), P(w[p + "n1.b"]), P(theta),
                 P(s["mem"]), P(s["memp"]), P(s["att"]), B, L, d, H, LN_EPS)
            # video self-attention
            for nm in ("q", "k", "v", "ao", "z2", "mem2", "z3"):
                sb(nm, (M, d))
            for nm in ("qT", "kT", "vT"):
                sb(nm, (B * d, Lp), fill=0)
            sb("lse", (B, H, Ls), f32, fill=float("inf"))
            gemm(fwd, p + "sa_q", s["memp"], w[p + "sa.wqk"][:d], w[p + "sa.bqk"][:d], out=s["q"], out_vt=s["qT"], vt_len=L, vt_pitch=Lp)
            gemm(fwd, p + "sa_k", s["memp"], w[p + "sa.wqk"][d:], w[p + "sa.bqk"][d:], out=s["k"], out_vt=s["kT"], vt_len=L, vt_pitch=Lp)
            gemm(fwd, p + "sa_v", s["mem"], w[p + "sa.wv"], w[p + "sa.bv"], out=s["v"], out_vt=s["vT"], vt_len=L, vt_pitch=Lp)
            attention(p + "sa_attn", s["q"], s["k"], s["vT"], s["ao"], s["lse"], L, L, Lp, Ls)
            gemm(fwd, p + "sa_out", s["ao"], w[p + "sa.wo"], w[p + "sa.bo"], out=s["z2"], residual=s["mem"])
            ln(p + "n2", s["z2"], w[p + "n2.w"], w[p + "n2.b"], s["mem2"])
            sb("pre1", (M, ff)); sb("hid1", (M, ff))
            gemm(fwd, p + "ffn1_up", s["mem2"], w[p + "mlp1.w1"], w[p + "mlp1.b1"], out=s["hid1"], act=ACT_GELU, out_pre=s["pre1"])
            gemm(fwd, p + "ffn1_down", s["hid1"], w[p + "mlp1.w2"], w[p + "mlp1.b2"], out=s["z3"], residual=s["mem2"])
            Xn, Xpn = buf(f"X{li + 1}", (M, d), bf), buf(f"Xp{li + 1}", (M, d), bf)
            layers_in_X.append(Xn); layers_in_Xp.append(Xpn)
            ln(p + "n3", s["z3"], w[p + "n3.w"], w[p + "n3.b"], Xn, y_pos=Xpn, theta_t=theta)
            # cross-attention K / V of the layer output
            for nm in ("kc", "vc"):
                sb(nm, (M, d))
            for nm in ("kcT", "vcT"):
                sb(nm, (B * d, Lp), fill=0)
            gemm(fwd, p + "ca_k", Xpn, w[p + "ca.wk"], w[p + "ca.bk"], out=s["kc"], out_vt=s["kcT"], vt_len=L, vt_pitch=Lp)
            gemm(fwd, p + "ca_v", Xn, w[p + "ca.wv"], w[p + "ca.bv"], out=s["vc"], out_vt=s["vcT"], vt_len=L, vt_pitch=Lp)
            # query self-attention
            for nm in ("qq", "kq", "vq", "aq", "z4", "o1", "o1p", "qc", "ac", "z5", "o2", "z6", "outp"):
                sb(nm, (MQ, d))
            for nm in ("qqT", "kqT", "vqT", "qcT"):
                sb(nm, (B * d, Qp), fill=0)
            sb("lse_q", (B, H, Qs), f32, fill=float("inf")); sb("lse_c", (B, H, Qs), f32, fill=float("inf"))
            s["out_in"], s["outp_in"] = out_cur, outp_cur
            q_first = len(fwd.calls)     # object-query branch: overlaps the frame-token chain of the next layer (engine.py:_Plan)
            gemm(fwd, p + "ta_q", outp_cur, w[p + "ta.wqk"][:d], w[p + "ta.bqk"][:d], out=s["qq"], out_vt=s["qqT"], vt_len=Q, vt_pitch=Qp)
            gemm(fwd, p + "ta_k", outp_cur, w[p + "ta.wqk"][d:], w[p + "ta.bqk"][d:], out=s["kq"], out_vt=s["kqT"], vt_len=Q, vt_pitch=Qp)
            gemm(fwd, p + "ta_v", out_cur, w[p + "ta.wv"], w[p + "ta.bv"], out=s["vq"], out_vt=s["vqT"], vt_len=Q, vt_pitch=Qp)
            attention(p + "ta_attn", s["qq"], s["kq"], s["vqT"], s["aq"], s["lse_q"], Q, Q, Qp, Qs)
            gemm(fwd, p + "ta_out", s["aq"], w[p + "ta.wo"], w[p + "ta.bo"], out=s["z4"], residual=out_cur)
            ln(p + "n4", s["z4"], w[p + "n4.w"], w[p + "n4.b"], s["o1"], y_pos=s["o1p"], pos_t=w["query_embed"], mod=Q)
            # query -> video cross-attention
            gemm(fwd, p + "ca_q", s["o1p"], w[p + "ca.wq"], w[p + "ca.bq"], out=s["qc"], out_vt=s["qcT"], vt_len=Q, vt_pitch=Qp)
            attention(p + "ca_attn", s["qc"], s["kc"], s["vcT"], s["ac"], s["lse_c"], Q, L, Lp, Qs, mask=vmask)
            gemm(fwd, p + "ca_out", s["ac"], w[p + "ca.wo"], w[p + "ca.bo"], out=s["z5"], residual=s["o1"])
            ln(p + "n5", s["z5"], w[p + "n5.w"], w[p + "n5.b"], s["o2"])
            sb("pre2", (MQ, ff)); sb("hid2", (MQ, ff))
            gemm(fwd, p + "ffn2_up", s["o2"], w[p + "mlp2.w1"], w[p + "mlp2.b1"], out=s["hid2"], act=ACT_GELU, out_pre=s["pre2"])
            gemm(fwd, p + "ffn2_down", s["hid2"], w[p + "mlp2.w2"], w[p + "mlp2.b2"], out=s["z6"], residual=s["o2"])
            ln(p + "n6", s["z6"], w[p + "n6.w"], w[p + "n6.b"], hs[li], y_pos=s["outp"], pos_t=w["query_embed"], mod=Q)
            for name, _, _ in fwd.calls[q_first:]:
                fwd.branch[name] = "q"
            fwd.after[p + "ca_attn"] = [p + "ca_v"]
            out_cur, outp_cur = hs[li], s["outp"]
            S.append(s)
        hs_all = hs.view(R, d)
        h1, h2 = buf("h1", (R, d), bf), buf("h2", (R, d), bf)
        logits, boxes = buf("logits", (NL, B, Q, 2), f32), buf("boxes", (NL, B, Q, 4), f32)
        gemm(fwd, "box0", hs_all, w["box.0.w"], w["box.0.b"], out=h1, act=ACT_RELU)
        gemm(fwd, "box1", h1, w["box.1.w"], w["box.1.b"], out=h2, act=ACT_RELU)
        call(fwd, "heads", lib.svol_heads, P(hs_all), P(h2), P(w["cls.w"]), P(w["cls.b"]), P(w["box.2.w"]), P(w["box.2.b"]), P(logits),
             P(boxes), R, d)
        for name in ("box0", "box1", "heads"):
            fwd.branch[name] = "q"

        # ================================================================= BACKWARD
        dlogits, dboxes = buf("dlogits", (NL, B, Q, 2), f32), buf("dboxes", (NL, B, Q, 4), f32)
        # scratch activation gradients
        gR = [buf(f"gR{i}", (R, d), bf) for i in range(3)]          # heads
        gq = [buf(f"gq{i}", (MQ, d), bf) for i in range(6)]
        gq_ff = buf("gq_ff", (MQ, ff), bf)
        gq_qk = buf("gq_qk", (MQ, 2 * d), bf)
        gv = [buf(f"gv{i}", (M, d), bf) for i in range(6)]
        gv_ff = buf("gv_ff", (M, ff), bf)
        gv_qk = buf("gv_qk", (M, 2 * d), bf)
        gv_in = buf("gv_in", (M, d_in), bf)
        # d(loss)/d(src_video): only when the caller's frame features require a gradient (a backbone trained through the
        # head, train.py:72): the input LayerNorm's backward then also writes dx (51 MB of bf16 at the headline shape)
        dx_in = buf("dx_in", (M, d_in), bf) if want_dx else None
        dOT_v = buf("dOT_v", (B * d, Lp), bf, fill=0)
        dOT_q = buf("dOT_q", (B * d, Qp), bf, fill=0)
        delta_v = buf("delta_v", (B, H, Ls), f32, fill=0)
        delta_q = buf("delta_q", (B, H, Qs), f32, fill=0)
        datt = buf("datt", (B, L), f32)
        dscores = buf("dscores", (B, H, L), f32)
        du = buf("du", (B, H, d), f32)
        dsk1, dsk0 = buf("dsk1", (B, d), f32), buf("dsk0", (B, d), f32)
        dsk_in = buf("dsk_in", (B, d_sk), f32)
        dhs_next = [buf(f"dhs_next{i}", (MQ, d), bf) for i in range(2)]    # d(out_in), d(outp_in) handed to the previous layer
        dX_next = buf("dX_next", (M, d), bf)                               # d(layer input X), handed to the previous layer
        # cross-attention dK / dV are produced by the query branch of layer i and consumed by its frame-token branch, which
        # runs concurrently with the query branch of layer i - 1: two alternating pairs
        dkc_b = [buf(f"dkc{i}", (M, d), bf) for i in range(2)]
        dvc_b = [buf(f"dvc{i}", (M, d), bf) for i in range(2)]

        be, ce = m.bbox_embed.layers, m.class_embed
        dhs_cls, dh2, dh1 = gR[0], gR[1], gR[2]
        bcall("heads_bwd", lib.svol_heads_backward, P(hs_all), P(h2), P(w["cls.w"]), P(w["box.2.w"]), P(boxes), P(dlogits), P(dboxes),
              P(dhs_cls), P(dh2), P(G(ce.weight)), P(G(ce.bias)), P(G(be[2].weight)), P(G(be[2].bias)), R, d)
        # dh2 already carries relu'(h2); box1: h2 = relu(h1 W1^T + b1)
        linear_bwd("box1", dh2, h1, G(be[1].weight), G(be[1].bias), wT=wt["box.1.wT"], dX=dh1, dact=h1, dact_mode=ACT_RELU)
        # dhs (from the box MLP) + dhs_cls  ->  gR[1] reused as the heads' total gradient w.r.t. hs_all
        dhs_heads = gR[1]
        linear_bwd("box0", dh1, hs_all, G(be[0].weight), G(be[0].bias), wT=wt["box.0.wT"], dX=dhs_heads, residual=dhs_cls)
        dqe = G(m.query_embed.weight)

        for li in reversed(range(NL)):
            layer = m.transformer.layers[li]
            p, s = f"l{li}.", S[li]
            X, Xp, Xn, Xpn = layers_in_X[li], layers_in_Xp[li], layers_in_X[li + 1], layers_in_Xp[li + 1]
            last = li == NL - 1
            ta, ca, sa = layer.token_self_attn, layer.content_token_cross_attn, layer.content_self_attn
            # ------------------------------------------------------------ query side
            bq_first = len(bwd.calls) if not last else 0     # (the last layer's range also covers the heads' backward)
            dhs_h = dhs_heads[li * MQ:(li + 1) * MQ]
            dz6, dhid, do2, dz5, dac = gq[0], gq_ff, gq[1], gq[2], gq[3]
            ln_bwd(p + "n6_bwd", s["z6"], dhs_h, w[p + "n6.w"], layer.norm6, dz6, dy2=None if last else dhs_next[0], dy3=None if last else dhs_next[1])
            if not last:       # outp = hs + query_embed is the q/k operand of the next layer's query self-attention
                bcall(p + "dqe_outp", lib.svol_batch_sum, P(dhs_next[1]), P(dqe), MQ, d, Q)
            linear_bwd(p + "ffn2_down", dz6, s["hid2"], G(layer.mlp2.fc2.weight), G(layer.mlp2.fc2.bias), wT=wt[p + "mlp2.w2T"], dX=dhid,
                       dact=s["pre2"], dact_mode=ACT_GELU)          # dhid = (dz6 W2) * gelu'(pre2)
            linear_bwd(p + "ffn2_up", dhid, s["o2"], G(layer.mlp2.fc1.weight), G(layer.mlp2.fc1.bias), wT=wt[p + "mlp2.w1T"], dX=do2,
                       residual=dz6)
            ln_bwd(p + "n5_bwd", s["z5"], do2, w[p + "n5.w"], layer.norm5, dz5)
            d_ = d
            linear_bwd(p + "ca_out", dz5, s["ac"], G(ca.out_proj.weight), G(ca.out_proj.bias), wT=wt[p + "ca.woT"], dX=dac,
                       out_vt=dOT_q, vt_len=Q, vt_pitch=Qp)
            dqc, dkc, dvc = gq[4], dkc_b[li & 1], dvc_b[li & 1]
            a = AttnBwdArgs()
            a.q, a.k, a.v, a.kt, a.qt, a.o, a.d_o, a.d_ot = P(s["qc"]), P(s["kc"]), P(s["vc"]), P(s["kcT"]), P(s["qcT"]), P(s["ac"]), P(dac), P(dOT_q)
            a.lse, a.delta, a.key_mask, a.dq, a.dk, a.dv = P(s["lse_c"]), P(delta_q), P(vmask), P(dqc), P(dkc), P(dvc)
            a.B, a.H, a.Lq, a.Lk = B, H, Q, L
            a.ldq = a.ldk = a.ldv = a.ld_o = a.ld_do = a.ld_dq = a.ld_dk = a.ld_dv = d
            a.kt_pitch, a.qt_pitch, a.stat_pitch = Lp, Qp, Qs
            bstruct(p + "ca_attn_bwd", lib.svol_attention_backward_bf16, a)
            do1p, do1 = gq[5], gq[3]
            gw, gb = G(ca.in_proj_weight), G(ca.in_proj_bias)
            linear_bwd(p + "ca_q", dqc, s["o1p"], gw[:d_], gb[:d_], wT=wt[p + "ca.wqT"], dX=do1p)
            bcall(p + "dqe_o1p", lib.svol_batch_sum, P(do1p), P(dqe), MQ, d, Q)
            dz4, daq = gq[0], gq[1]
            ln_bwd(p + "n4_bwd", s["z4"], dz5, w[p + "n4.w"], layer.norm4, dz4, dy2=do1p)        # o1 feeds the residual (dz5) and o1 + qe (do1p)
            linear_bwd(p + "ta_out", dz4, s["aq"], G(ta.out_proj.weight), G(ta.out_proj.bias), wT=wt[p + "ta.woT"], dX=daq,
                       out_vt=dOT_q, vt_len=Q, vt_pitch=Qp)
            dvq = gq[2]
            a = AttnBwdArgs()
            a.q, a.k, a.v, a.kt, a.qt, a.o, a.d_o, a.d_ot = P(s["qq"]), P(s["kq"]), P(s["vq"]), P(s["kqT"]), P(s["qqT"]), P(s["aq"]), P(daq), P(dOT_q)
            a.lse, a.delta, a.key_mask = P(s["lse_q"]), P(delta_q), None
            a.dq, a.dk, a.dv = P(gq_qk), P(gq_qk[:, d:]), P(dvq)
            a.B, a.H, a.Lq, a.Lk = B, H, Q, Q
            a.ldq = a.ldk = a.ldv = a.ld_o = a.ld_do = a.ld_dv = d
            a.ld_dq = a.ld_dk = 2 * d
            a.kt_pitch, a.qt_pitch, a.stat_pitch = Qp, Qp, Qs
            bstruct(p + "ta_attn_bwd", lib.svol_attention_backward_bf16, a)
            gw, gb = G(ta.in_proj_weight), G(ta.in_proj_bias)
            # d(outp_in) = [dq | dk] [Wq ; Wk];  d(out_in) = dv Wv + dz4 (residual of norm4's input)
            linear_bwd(p + "ta_qk", gq_qk, s["outp_in"], gw[:2 * d_], gb[:2 * d_], wT=wt[p + "ta.wqkT"], dX=dhs_next[1])
            linear_bwd(p + "ta_v", dvq, s["out_in"], gw[2 * d_:], gb[2 * d_:], wT=wt[p + "ta.wvT"], dX=dhs_next[0], residual=dz4)
            if li == 0:        # out_in = 0 (no parameters behind it); outp_in = query_embed broadcast
                bcall(p + "dqe_qe", lib.svol_batch_sum, P(dhs_next[1]), P(dqe), MQ, d, Q)
            # Branches of the backward graph: the query side of layer i - 1 only needs the query side of layer i, so it runs
            # concurrently with the frame-token side of layer i (which waits for layer i's cross-attention backward).
            for name, _, _ in bwd.calls[bq_first:]:
                bwd.branch[name] = "q"
            bv_first = len(bwd.calls)
            # ------------------------------------------------------------ frame-token side
            gw, gb = G(ca.in_proj_weight), G(ca.in_proj_bias)
            dXpn, dXn = gv[2], gv[3]
            linear_bwd(p + "ca_k", dkc, Xpn, gw[d_:2 * d_], gb[d_:2 * d_], wT=wt[p + "ca.wkT"], dX=dXpn)
            linear_bwd(p + "ca_v", dvc, Xn, gw[2 * d_:], gb[2 * d_:], wT=wt[p + "ca.wvT"], dX=dXn, residual=dXpn)
            bwd.after[bwd.calls[bv_first][0]] = [p + "ca_attn_bwd"]
            dz3 = gv[0]
            ln_bwd(p + "n3_bwd", s["z3"], dXn, w[p + "n3.w"], layer.norm3, dz3, dy2=None if last else dX_next)
            dmem2 = gv[1]
            linear_bwd(p + "ffn1_down", dz3, s["hid1"], G(layer.mlp1.fc2.weight), G(layer.mlp1.fc2.bias), wT=wt[p + "mlp1.w2T"], dX=gv_ff,
                       dact=s["pre1"], dact_mode=ACT_GELU)
            linear_bwd(p + "ffn1_up", gv_ff, s["mem2"], G(layer.mlp1.fc1.weight), G(layer.mlp1.fc1.bias), wT=wt[p + "mlp1.w1T"], dX=dmem2,
                       residual=dz3)
            dz2, dao = gv[2], gv[3]
            ln_bwd(p + "n2_bwd", s["z2"], dmem2, w[p + "n2.w"], layer.norm2, dz2)
            linear_bwd(p + "sa_out", dz2, s["ao"], G(sa.out_proj.weight), G(sa.out_proj.bias), wT=wt[p + "sa.woT"], dX=dao, out_vt=dOT_v,
                       vt_len=L, vt_pitch=Lp)
            dv_ = gv[0]
            a = AttnBwdArgs()
            a.q, a.k, a.v, a.kt, a.qt, a.o, a.d_o, a.d_ot = P(s["q"]), P(s["k"]), P(s["v"]), P(s["kT"]), P(s["qT"]), P(s["ao"]), P(dao), P(dOT_v)
            a.lse, a.delta, a.key_mask = P(s["lse"]), P(delta_v), None
            a.dq, a.dk, a.dv = P(gv_qk), P(gv_qk[:, d:]), P(dv_)
            a.B, a.H, a.Lq, a.Lk = B, H, L, L
            a.ldq = a.ldk = a.ldv = a.ld_o = a.ld_do = a.ld_dv = d
            a.ld_dq = a.ld_dk = 2 * d
            a.kt_pitch, a.qt_pitch, a.stat_pitch = Lp, Lp, Ls
            bstruct(p + "sa_attn_bwd", lib.svol_attention_backward_bf16, a)
            gw, gb = G(sa.in_proj_weight), G(sa.in_proj_bias)
            dmemp, dmem = gv[1], gv[3]
            linear_bwd(p + "sa_qk", gv_qk, s["memp"], gw[:2 * d_], gb[:2 * d_], wT=wt[p + "sa.wqkT"], dX=dmemp)
            linear_bwd(p + "sa_v", dv_, s["mem"], gw[2 * d_:], gb[2 * d_:], wT=wt[p + "sa.wvT"], dX=dmem, residual=dmemp)
            # norm1 over x * (1 + att): dx = dz (1 + att), datt; mem feeds the residual (dz2), memp/mem (dmem)
            dxg = gv[0]
            ln_bwd(p + "n1_bwd", X, dmem, w[p + "n1.w"], layer.norm1, dxg, dy2=dz2, att=s["att"], datt=datt)
            bcall(p + "gate_bwd", lib.svol_gate_backward, P(Xp), P(s["u"]), P(s["scores"]), P(datt), P(dxg), P(dX_next), P(dscores), P(du),
                  B, L, d, H)
            g_ = layer.sketch_video_cross_attn
            bcall(p + "gate_vec_bwd", lib.svol_gate_vectors_backward, P(sk1), P(w[p + "gate.w"]), P(w[p + "gate.b"]), P(du),
                  P(G(g_.in_proj_weight)), P(G(g_.in_proj_bias)), P(dsk1), B, d, H)

        # ---- input projection of the frame tokens: X0 = xn1 W1^T + b1; xn1 = LN(a0); a0 = relu(xn0 W0^T + b0); xn0 = LN(x_in)
        dxn1, da0 = gv[1], gv[2]
        linear_bwd("in_proj1", dX_next, xn1, G(vp[1].net[1].weight), G(vp[1].net[1].bias), wT=wt["in_video.1.wT"], dX=dxn1)
        ln_bwd("in_ln1_bwd", a0, dxn1, w["in_video.1.ln_w"], vp[1].LayerNorm, da0, site=1)
        bcall("in_relu_bwd", lib.svol_act_backward, P(da0), P(a0), P(da0), M * d, ACT_RELU)
        linear_bwd("in_proj0", da0, xn0, G(vp[0].net[1].weight), G(vp[0].net[1].bias), wT=wt["in_video.0.wT"], dX=gv_in)
        ln_bwd("in_ln0_bwd", x_in, gv_in, w["in_video.0.ln_w"], vp[0].LayerNorm, dx_in, z_f32=True, cols=d_in, site=0)
        # ---- sketch branch
        bcall("sk_proj1_bwd", lib.svol_ln_linear_f32_backward, P(sk0), P(w["in_sketch.1.ln_w"]), P(w["in_sketch.1.ln_b"]),
              P(w["in_sketch.1.w"]), P(sk1), P(dsk1), 0, P(dsk0), P(G(sp[1].LayerNorm.weight)), P(G(sp[1].LayerNorm.bias)),
              P(G(sp[1].net[1].weight)), P(G(sp[1].net[1].bias)), B, d, d, LN_EPS, drop_p, P(seed), 3)
        bcall("sk_proj0_bwd", lib.svol_ln_linear_f32_backward, P(s_in), P(w["in_sketch.0.ln_w"]), P(w["in_sketch.0.ln_b"]),
              P(w["in_sketch.0.w"]), P(sk0), P(dsk0), 1, P(dsk_in), P(G(sp[0].LayerNorm.weight)), P(G(sp[0].LayerNorm.bias)),
              P(G(sp[0].net[1].weight)), P(G(sp[0].net[1].bias)), B, d_sk, d, LN_EPS, drop_p, P(seed), 2)
        return {"fwd": fwd, "bwd": bwd, "buf": bufs}

    def plan_for(self, B: int, L: int, d_in: int, want_dx: bool = False) -> dict:
        self._setup_grads()
        self._pack_transposed()
        key = (B, L, d_in, bool(want_dx))
        plan = self._plans.get(key)
        if plan is None:
            plan = self._build(B, L, d_in, bool(want_dx))
            self._plans[key] = plan
        return plan

    # ------------------------------------------------------------------ run
    @torch.no_grad()
    def forward(self, src_sketch, src_sketch_mask, src_video, src_video_mask, want_input_grads: bool = False):
        """Training forward.  Returns (logits [NL,B,Q,2], boxes [NL,B,Q,4]) fp32 views of the plan's buffers.
        ``want_input_grads``: the backward also produces d/d(src_video) (``input_grads()``)."""
        _lib.require_device()
        B, L, d_in = src_video.shape
        plan = self.plan_for(B, L, d_in, want_input_grads)
        b = plan["buf"]
        b["src_video"].copy_(src_video, non_blocking=True)
        if src_sketch.dim() == 3 and src_sketch.shape[1] != 1:
            raise NotImplementedError("svol_b200 supports one sketch token per pair (L_sketch == 1)")
        b["src_sketch"].copy_(src_sketch.reshape(B, -1), non_blocking=True)
        b["src_video_mask"].copy_(src_video_mask, non_blocking=True)
        self.last_seed = self.seed_base + self.step_index      # this step's dropout stream; the backward reuses it
        self.step_index += 1
        b["seed"].fill_(self.last_seed)
        self._run(plan, "fwd")
        self._last = plan
        self.forward_token += 1          # identifies whose activations the plan buffers hold (see backward)
        return b["logits"], b["boxes"]

    def input_grads(self):
        """(d/d src_sketch [B, 1, D_s] fp32, d/d src_video [B, L, D_v] fp32 or None) of the most recent backward, as new
        tensors.  The frame-feature gradient exists when the forward was run with ``want_input_grads``; like every
        activation gradient of this backward it is computed in bf16."""
        b = self._last["buf"]
        B = b["src_sketch"].shape[0]
        g_sk = b["dsk_in"].view(B, 1, -1).clone()
        g_v = b["dx_in"].view(b["src_video"].shape).float() if "dx_in" in b else None
        return g_sk, g_v

    def _run(self, plan: dict, which: str) -> None:
        """Eager on the first call of a shape (module loading, function attributes), captured on the second,
        replayed afterwards.  The backward graph includes the zeroing of the gradient buffers."""
        def body(side=None):
            if which == "bwd":
                self.grad_flat.zero_()
                plan["buf"]["dsk1"].zero_()
            plan[which].run(torch.cuda.current_stream().cuda_stream, side=side)

        if not self.use_graph:
            return body()
        state = plan.setdefault("graphs", {})
        if which not in state:
            state[which] = None                      # warm-up run
            return body()
        if state[which] is None:
            torch.cuda.synchronize()
            if self._side is None:
                self._side = (torch.cuda.Stream(), torch.cuda.Stream())
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                body(side=self._side)            # the query / frame-token branches become concurrent graph branches
            state[which] = g
        state[which].replay()

    @torch.no_grad()
    def backward(self, grad_logits: torch.Tensor, grad_boxes: torch.Tensor, token: int = None) -> None:
        """Runs the backward of the most recent training forward; parameter gradients land in ``grad_of(p)``.
        ``token`` (the value of ``forward_token`` right after the forward being differentiated) guards against a second
        forward having overwritten the saved activations (they live in static plan buffers, one set per shape)."""
        if token is not None and token != self.forward_token:
            raise RuntimeError("svol_b200: backward() of a training forward whose saved activations were overwritten by a later "
                               "forward; call loss.backward() before the next model(...) call (one forward in flight)")
        plan = self._last
        b = plan["buf"]
        b["dlogits"].copy_(grad_logits.reshape(b["dlogits"].shape))
        b["dboxes"].copy_(grad_boxes.reshape(b["dboxes"].shape))
        self._run(plan, "bwd")
