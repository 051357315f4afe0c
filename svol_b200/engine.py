"""Launch plan of the SVANet head forward on the CUDA C ABI.

``HeadEngine`` turns one ``SVANet`` module (reference-compatible parameters, see
``svol_b200/modeling/svanet.py``) into a static sequence of ``libsvol_b200.so`` calls:

  * weights are packed once into engine-owned bf16 / fp32 buffers (q rows pre-scaled by
    log2(e)/sqrt(d_head) so the attention softmax is a bare ex2) and re-packed when a parameter's
    version counter or storage changes;
  * activations live in a per-(batch, video length) workspace whose addresses never change, so the
    forward after its first kernel is captured in one CUDA graph and replayed (``use_graph=True``, the
    default; the first kernel reads the caller's frame features in place);
  * activations are batch-major ``[B*L, 256]`` bf16 row matrices; attention V operands are stored
    transposed per head (``[B*8*32, L_pad]``) directly by the projection GEMM's epilogue.

PyTorch is used for memory, streams and graph capture only -- every arithmetic step of the forward
is a kernel of this repository.  Reference call order: ``SVANet.forward`` (lib/modeling/svanet.py:65-141)
-> ``CrossModalTransformer.forward`` (lib/modeling/cross_modal_transformer.py:27-81) ->
``CrossModalTransformerLayer.forward`` (:105-160).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, List, Tuple

import torch

from . import _lib
from ._lib import ACT_GELU, ACT_NONE, ACT_RELU, AttnArgs, FfnArgs, GemmArgs

LN_EPS = 1e-5
HEAD_DIM = 32


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class _Plan:
    """A recorded list of C-ABI calls plus the buffers they touch.

    ``calls`` is a valid sequential order.  Every call also carries a branch tag -- 'v' (frame-token chain, the
    launching stream), 'q' (object-query chain) or 'in' (kernels that depend on the inputs only: sketch branch, gate
    vectors) -- and the names of the calls on OTHER branches it must wait for.  Replayed as a CUDA graph the three
    branches run concurrently: in a decoder layer the query self-attention block (cross_modal_transformer.py:145-149)
    does not depend on that layer's frame-token block (:122-143), only the cross-attention (:151-156) joins them, so
    the small, launch-latency-bound query kernels (80 CTAs on 148 SMs) fill the tails of the large frame-token
    kernels instead of extending the critical path."""

    def __init__(self):
        self.calls: List[Tuple[str, object, tuple]] = []
        self.branch: Dict[str, str] = {}
        self.after: Dict[str, List[str]] = {}
        self.ln_in_index = 0
        self.keep: List[object] = []        # ctypes structs must outlive the plan
        self.buf: Dict[str, torch.Tensor] = {}
        self.graph = None
        self.graph_full = None              # same, including the input LayerNorm (inputs written into the plan's own buffers)

    def run(self, stream: int, start: int = 0, side=None) -> None:
        """Launches calls[start:].  ``side`` = None: sequentially on ``stream``.  ``side`` = (query stream, input
        stream): each call on its branch's stream with event edges for the cross-branch dependencies (fork at the
        start, join at the end) -- used under CUDA-graph capture."""
        if side is None:
            for name, fn, args in self.calls[start:]:
                rc = fn(*args, stream)
                if rc != 0:
                    _lib.check(rc, name)
            return
        main = torch.cuda.current_stream()
        assert main.cuda_stream == stream
        streams = {"v": main, "q": side[0], "in": side[1]}
        for st in side:
            st.wait_stream(main)                                   # fork
        launched = {name for name, _, _ in self.calls[:start]}
        needed = {dep for deps in self.after.values() for dep in deps}
        events: Dict[str, torch.cuda.Event] = {}
        for name, fn, args in self.calls[start:]:
            st = streams[self.branch.get(name, "v")]
            for dep in self.after.get(name, ()):
                if dep in launched and dep not in events:
                    continue                                       # ran before the fork: already ordered
                if self.branch.get(dep, "v") != self.branch.get(name, "v"):
                    st.wait_event(events[dep])
            _lib.check(fn(*args, st.cuda_stream), name)
            if name in needed:
                ev = torch.cuda.Event()
                ev.record(st)
                events[name] = ev
        for st in side:
            main.wait_stream(st)                                   # join


class WeightPacker:
    """Keeps the packed operand copies of a module's parameters (bf16 GEMM operands, scaled / sliced / transposed
    views, fp32 biases) in persistent buffers and refreshes ALL of them with one ``svol_pack_weights`` launch.

    ``put(name, param, dtype, row0, rows, scaled_rows, scale, transpose)`` declares ``dst = param[row0:row0+rows]``
    (rows of a [out, in] weight, or elements of a bias) with the first ``scaled_rows`` rows multiplied by ``scale``,
    cast to ``dtype`` and optionally transposed.  Buffers keep their addresses across refreshes (plans and CUDA graphs
    stay valid); the job table is re-uploaded only when a parameter's storage moved."""

    def __init__(self):
        self.tensors: Dict[str, torch.Tensor] = {}
        self._jobs: List[tuple] = []
        self._table = None
        self._table_key = None
        self._moved = False
        self._dev = None

    def begin(self, device):
        self._jobs, self._moved, self._dev = [], False, device

    def put(self, name, param, dtype, row0=0, rows=None, scaled_rows=0, scale=1.0, transpose=False):
        t = param.detach()
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise NotImplementedError("svol_b200 packs contiguous fp32 parameters")
        cols = t.shape[1] if t.dim() == 2 else 1
        total = t.shape[0]
        rows = total - row0 if rows is None else rows
        shape = ((cols, rows) if transpose else (rows, cols)) if t.dim() == 2 else (rows,)
        old = self.tensors.get(name)
        if old is None or tuple(old.shape) != shape or old.dtype != dtype or old.device != t.device:
            self.tensors[name] = torch.empty(shape, device=t.device, dtype=dtype)
            self._moved = True
        flags = (_lib.PACK_BF16 if dtype == torch.bfloat16 else 0) | (_lib.PACK_TRANSPOSE if transpose else 0)
        self._jobs.append((t.data_ptr() + row0 * cols * 4, self.tensors[name].data_ptr(), rows, cols, scaled_rows, flags, float(scale)))

    def end(self) -> bool:
        """Uploads the job table if it changed and runs the packing launch.  Returns True if any buffer moved."""
        key = tuple(self._jobs)
        if key != self._table_key:
            arr = (_lib.PackJob * len(self._jobs))()
            for i, (src, dst, rows, cols, sr, flags, scale) in enumerate(self._jobs):
                arr[i].src, arr[i].dst, arr[i].rows, arr[i].cols = src, dst, rows, cols
                arr[i].scaled_rows, arr[i].flags, arr[i].scale = sr, flags, scale
            host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
            self._table = host.to(self._dev)
            self._table_key = key
        if self._dev.type != "cuda":
            raise RuntimeError("svol_b200 packs weights on a CUDA device; there is no CPU fallback")
        _lib.check(_lib.get_lib().svol_pack_weights(self._table.data_ptr(), len(self._jobs), _lib.stream_ptr()), "pack_weights")
        return self._moved


class HeadEngine:
    def __init__(self, module, use_graph: bool = True):
        self.module = module
        self.use_graph = use_graph
        self.plain = os.environ.get("SVOL_B200_PLAIN", "0") == "1"   # debug: SIMT kernels instead of tcgen05
        self._w: Dict[str, torch.Tensor] = {}
        self._wstate = None
        self._plist = None
        self.weights_generation = 0        # bumped by every repack: consumers of derived copies (TrainEngine's W^T) compare it
        self._packer = WeightPacker()
        self._plans: Dict[Tuple[int, int, int, int], _Plan] = {}
        self._side = None                  # (query, input) capture streams: the forked branches of the forward graph
        self.launches_per_forward = 0

    # ------------------------------------------------------------------ weights
    def _param_state(self):
        """Cheap change signature of the module's parameters (this runs on every forward: walking module.parameters() and
        calling data_ptr() on each of the ~100 tensors cost 0.17 ms per step, most of the host's enqueue time).  In-place
        updates (optimizers, load_state_dict) bump ``_version``; ``.to()`` / ``.cuda()`` move every storage, the first and the
        last included; FusedAdamW, which updates through a raw pointer, resets ``_wstate`` explicitly.  The Parameter list
        itself is re-read whenever the state was reset."""
        if self._wstate is None or self._plist is None:
            self._plist = list(self.module.parameters())
        ps = self._plist
        return (ps[0].data_ptr(), ps[-1].data_ptr(), tuple([p._version for p in ps]))

    @torch.no_grad()
    def _pack_weights(self) -> None:
        m = self.module
        d = m.transformer.d_model
        if d != 256 or m.transformer.nhead != 8:
            raise NotImplementedError("svol_b200 kernels are built for hidden_dim 256 / 8 heads (head_dim 32)")
        qscale = math.log2(math.e) / math.sqrt(HEAD_DIM)
        pk = self._packer
        pk.begin(next(m.parameters()).device)
        put = pk.put
        bf, f32 = torch.bfloat16, torch.float32
        for which in ("video", "sketch"):
            seq = getattr(m, f"input_{which}_proj")
            for i, lin in enumerate(seq):
                put(f"in_{which}.{i}.ln_w", lin.LayerNorm.weight, f32)
                put(f"in_{which}.{i}.ln_b", lin.LayerNorm.bias, f32)
                put(f"in_{which}.{i}.w", lin.net[1].weight, bf if which == "video" else f32)
                put(f"in_{which}.{i}.b", lin.net[1].bias, f32)
        put("query_embed", m.query_embed.weight, f32)
        for li, layer in enumerate(m.transformer.layers):
            p = f"l{li}."
            g = layer.sketch_video_cross_attn
            put(p + "gate.w", g.in_proj_weight, f32)
            put(p + "gate.b", g.in_proj_bias, f32)
            for tag, att in (("sa", layer.content_self_attn), ("ta", layer.token_self_attn)):
                W, b = att.in_proj_weight, att.in_proj_bias
                # rows [0, d) = q (pre-scaled by log2(e)/sqrt(d_head): the attention softmax is a bare ex2), [d, 2d) = k
                put(p + tag + ".wqk", W, bf, rows=2 * d, scaled_rows=d, scale=qscale)
                put(p + tag + ".bqk", b, f32, rows=2 * d, scaled_rows=d, scale=qscale)
                put(p + tag + ".wv", W, bf, row0=2 * d)
                put(p + tag + ".bv", b, f32, row0=2 * d)
                put(p + tag + ".wo", att.out_proj.weight, bf)
                put(p + tag + ".bo", att.out_proj.bias, f32)
                # one launch computes q | k from x + pos and v from x (split GEMM): [q * scale ; k ; v]
                put(p + tag + ".wqkv", W, bf, scaled_rows=d, scale=qscale)
                put(p + tag + ".bqkv", b, f32, scaled_rows=d, scale=qscale)
            ca = layer.content_token_cross_attn
            W, b = ca.in_proj_weight, ca.in_proj_bias
            put(p + "ca.wq", W, bf, rows=d, scaled_rows=d, scale=qscale)
            put(p + "ca.bq", b, f32, rows=d, scaled_rows=d, scale=qscale)
            put(p + "ca.wk", W, bf, row0=d, rows=d)
            put(p + "ca.bk", b, f32, row0=d, rows=d)
            put(p + "ca.wv", W, bf, row0=2 * d)
            put(p + "ca.bv", b, f32, row0=2 * d)
            put(p + "ca.wo", ca.out_proj.weight, bf)
            put(p + "ca.bo", ca.out_proj.bias, f32)
            put(p + "ca.wkv", W, bf, row0=d)             # k from mem + pos, v from mem: one split launch
            put(p + "ca.bkv", b, f32, row0=d)
            for n in range(1, 7):
                norm = getattr(layer, f"norm{n}")
                put(p + f"n{n}.w", norm.weight, f32)
                put(p + f"n{n}.b", norm.bias, f32)
            for tag, mlp in (("mlp1", layer.mlp1), ("mlp2", layer.mlp2)):
                put(p + tag + ".w1", mlp.fc1.weight, bf)
                put(p + tag + ".b1", mlp.fc1.bias, f32)
                put(p + tag + ".w2", mlp.fc2.weight, bf)
                put(p + tag + ".b2", mlp.fc2.bias, f32)
        for i in range(2):
            put(f"box.{i}.w", m.bbox_embed.layers[i].weight, bf)
            put(f"box.{i}.b", m.bbox_embed.layers[i].bias, f32)
        put("box.2.w", m.bbox_embed.layers[2].weight, f32)
        put("box.2.b", m.bbox_embed.layers[2].bias, f32)
        put("cls.w", m.class_embed.weight, f32)
        put("cls.b", m.class_embed.bias, f32)
        if pk.end():
            self._plans.clear()                 # a packed buffer moved: recorded plans hold stale addresses
        self._w = pk.tensors
        self.weights_generation += 1

    def _weights(self) -> Dict[str, torch.Tensor]:
        st = self._param_state()
        if st != self._wstate:
            self._pack_weights()
            self._wstate = st
        return self._w

    # ------------------------------------------------------------------ plan
    def _build_plan(self, B: int, L: int, d_in: int) -> _Plan:
        m, w = self.module, self._w
        lib = _lib.get_lib()
        dev = w["cls.w"].device
        d, H, ff = 256, 8, m.transformer.layers[0].mlp1.fc1.weight.shape[0]
        Q = m.num_queries
        NL = len(m.transformer.layers)
        n_proj = len(m.input_video_proj)
        d_sk = m.input_sketch_proj[0].net[1].weight.shape[1]
        M, MQ = B * L, B * Q
        Lp, Qp = _round_up(L, 8), _round_up(Q, 8)
        plan = _Plan()
        bf, f32 = torch.bfloat16, torch.float32

        def buf(name, shape, dtype, zero=False):
            t = (torch.zeros if zero else torch.empty)(shape, device=dev, dtype=dtype)
            plan.buf[name] = t
            return t

        # inputs (static addresses; HeadEngine.forward copies the caller's tensors in)
        x_in = buf("src_video", (B, L, d_in), f32)
        s_in = buf("src_sketch", (B, d_sk), f32)
        vmask = buf("src_video_mask", (B, L), f32)
        # activations
        xn = buf("xn", (M, d_in), bf)
        ha, hb = buf("ha", (M, d), bf), buf("hb", (M, d), bf)
        # sine positions: the tcgen05 path evaluates them inside the consuming epilogues from one angle per token;
        # the 1 KB/token fp32 table is only materialised for the SIMT debug path
        pos = buf("pos", (M, d), f32) if self.plain else None
        theta = buf("theta", (M,), f32)
        X, Xp = buf("X", (M, d), bf), buf("Xp", (M, d), bf)
        mem, memp = buf("mem", (M, d), bf), buf("memp", (M, d), bf)
        mem2 = buf("mem2", (M, d), bf)
        qk = buf("qk", (M, 2 * d), bf)
        vt = buf("vt", (B * d, Lp), bf, zero=True)
        # cross-attention K / V^T and the gate vectors are per layer: the frame-token branch of layer i+1 may be
        # producing them while the query branch is still consuming layer i's
        vt_ca = [buf(f"vt_ca{li}", (B * d, Lp), bf, zero=True) for li in range(NL)]
        att = buf("att", (M, d), bf)
        sk_a, sk_b = buf("sk_a", (B, d), f32), buf("sk_b", (B, d), f32)
        u = [buf(f"u{li}", (B, H, d), f32) for li in range(NL)]
        scores = buf("scores", (B, H, L), f32)
        zeros_q = buf("zeros_q", (MQ, d), bf, zero=True)
        qe_bf = buf("qe_bf", (MQ, d), bf)
        o1, o1p = buf("o1", (MQ, d), bf), buf("o1p", (MQ, d), bf)
        o2 = buf("o2", (MQ, d), bf)
        outp = buf("outp", (MQ, d), bf)
        qkq = buf("qkq", (MQ, 2 * d), bf)
        vtq = buf("vtq", (B * d, Qp), bf, zero=True)
        attq = buf("attq", (MQ, d), bf)
        qc = buf("qc", (MQ, d), bf)
        kc = [buf(f"kc{li}", (M, d), bf) for li in range(NL)]
        hs = buf("hs", (NL, MQ, d), bf)
        h1, h2 = buf("h1", (NL * MQ, d), bf), buf("h2", (NL * MQ, d), bf)
        logits = buf("logits", (NL, B, Q, 2), f32)
        boxes = buf("boxes", (NL, B, Q, 4), f32)
        logits._svol_static = boxes._svol_static = True       # plan-owned outputs: stable addresses (criterion graph replay)

        P = _lib.ptr
        gemm_fn = lib.svol_gemm_bf16_plain if self.plain else lib.svol_gemm_bf16
        attn_fn = lib.svol_attention_bf16_plain if self.plain else lib.svol_attention_bf16

        def gemm(name, A, W, bias, out=None, act=ACT_NONE, residual=None, ln=None, out_pos=None, pos_t=None,
                 pos_mod=0, out_vt=None, vt_len=0, vt_pitch=0, theta_t=None, A2=None, split=0):
            a = GemmArgs()
            a.A, a.W = P(A), P(W)
            a.M, a.K = A.shape
            a.N = W.shape[0]
            assert W.shape[1] == a.K
            a.lda, a.ldw = A.stride(0), W.stride(0)
            e = a.ep
            e.bias, e.act = P(bias), act
            e.residual, e.ld_res = P(residual), (residual.stride(0) if residual is not None else 0)
            if ln is not None:
                e.ln_weight, e.ln_bias, e.ln_eps = P(ln[0]), P(ln[1]), LN_EPS
            e.out = P(out)
            e.out_pos = P(out_pos)
            ld_o = out.stride(0) if out is not None else (out_pos.stride(0) if out_pos is not None else 0)
            e.ld_out = ld_o
            if out_pos is not None and theta_t is not None:
                e.pos_theta = P(theta_t)
            elif out_pos is not None:
                e.pos, e.ld_pos, e.pos_row_mod = P(pos_t), pos_t.stride(0), pos_mod
            if out_vt is not None:
                e.out_vt, e.vt_len, e.vt_pitch = P(out_vt), vt_len, vt_pitch
            if A2 is not None:
                a.A2, a.lda2, a.split_block = P(A2), A2.stride(0), split
            plan.keep.append(a)
            plan.calls.append((name, gemm_fn, (C.byref(a),)))

        def ffn(name, x, w1, b1, w2, b2, ln, out, out_pos=None, pos_t=None, pos_mod=0, theta_t=None):
            a = FfnArgs()
            a.x, a.w1, a.b1, a.w2, a.b2 = P(x), P(w1), P(b1), P(w2), P(b2)
            a.ln_weight, a.ln_bias, a.out, a.out_pos = P(ln[0]), P(ln[1]), P(out), P(out_pos)
            a.pos, a.pos_theta = P(pos_t), P(theta_t)
            a.M, a.d, a.ff = x.shape[0], x.shape[1], w1.shape[0]
            a.ldx, a.ldw1, a.ldw2, a.ld_out = x.stride(0), w1.stride(0), w2.stride(0), out.stride(0)
            a.ld_pos, a.pos_row_mod = (pos_t.stride(0) if pos_t is not None else 0), pos_mod
            a.ln_eps = LN_EPS
            plan.keep.append(a)
            plan.calls.append((name, lib.svol_ffn_bf16, (C.byref(a),)))

        def attention(name, q, k, vt_t, out, Lq, Lk, ldq, ldk, pitch, mask=None):
            a = AttnArgs()
            a.q, a.k, a.vt, a.key_mask, a.out = P(q), P(k), P(vt_t), P(mask), P(out)
            a.B, a.H, a.Lq, a.Lk, a.ldq, a.ldk, a.ldo, a.vt_pitch = B, H, Lq, Lk, ldq, ldk, out.stride(0), pitch
            plan.keep.append(a)
            plan.calls.append((name, attn_fn, (C.byref(a),)))

        def call(name, fn, *args):
            plan.calls.append((name, fn, args))

        def tag(branch, after=()):
            """Branch / cross-branch dependencies of the most recently added call."""
            name = plan.calls[-1][0]
            plan.branch[name] = branch
            if after:
                plan.after[name] = list(after)

        # the gate runs as one cluster kernel when a sample's token rows fit one cluster's shared memory (L <= 3384);
        # longer clips (config C4) and the SIMT debug path keep gate_scores + gate_apply (SVOL_B200_GATE=split forces them)
        fused_gate = (not self.plain and os.environ.get("SVOL_B200_GATE", "fused") != "split"
                      and bool(lib.svol_gate_fused_supported(L)))
        # ---- input projection of the frame tokens (svanet.py:49-55,83): LN -> Linear -> ReLU -> LN -> Linear
        plan.ln_in_index = len(plan.calls)
        call("ln_in", lib.svol_layernorm_f32_to_bf16, P(x_in), P(w["in_video.0.ln_w"]), P(w["in_video.0.ln_b"]), P(xn),
             M, d_in, LN_EPS)
        if self.plain:
            call("posenc", lib.svol_posenc_sine, P(vmask), P(pos), B, L, d)
        call("posenc_theta", lib.svol_posenc_theta, P(vmask), P(theta), B, L)
        cur = xn
        for i in range(n_proj):
            last = i == n_proj - 1
            dst = X if last else (ha if cur is not ha else hb)
            gemm(f"in_proj{i}", cur, w[f"in_video.{i}.w"], w[f"in_video.{i}.b"], out=dst,
                 act=ACT_NONE if last else ACT_RELU,
                 ln=None if last else (w[f"in_video.{i + 1}.ln_w"], w[f"in_video.{i + 1}.ln_b"]),
                 out_pos=Xp if (last and not fused_gate) else None, pos_t=pos if last else None,   # x + pos only feeds gate_scores
                 theta_t=theta if (last and not self.plain) else None)
            cur = dst
        # ---- sketch branch (svanet.py:56-60,87), fp32, B rows
        s_cur = s_in
        for i in range(n_proj):
            s_dst = sk_a if s_cur is not sk_a else sk_b
            dim_in = d_sk if i == 0 else d
            call(f"sk_proj{i}", lib.svol_ln_linear_f32, P(s_cur), P(w[f"in_sketch.{i}.ln_w"]), P(w[f"in_sketch.{i}.ln_b"]),
                 P(w[f"in_sketch.{i}.w"]), P(w[f"in_sketch.{i}.b"]), 0 if i == n_proj - 1 else 1, P(s_dst), B, dim_in, d,
                 LN_EPS)
            tag("in")
            s_cur = s_dst
        # ---- query side: out0 = 0, q/k operand = query_embed broadcast (cross_modal_transformer.py:52-56)
        call("qe_bcast", lib.svol_add_pos_bf16, P(w["query_embed"]), None, P(qe_bf), MQ, d, Q)
        tag("q")
        out_cur, outp_cur = zeros_q, qe_bf
        x_cur, xp_cur = X, Xp
        for li in range(NL):
            p = f"l{li}."
            # (a) sketch-conditioned gate + norm1                                        :122-127
            call(p + "gate_vec", lib.svol_gate_vectors, P(s_cur), P(w[p + "gate.w"]), P(w[p + "gate.b"]), P(u[li]), B, d, H)
            tag("in")
            if fused_gate:
                # one launch: scores from x + pos (fp32), softmax over the sample closed inside a cluster, norm1
                call(p + "gate_fused", lib.svol_gate_fused, P(x_cur), P(u[li]), P(w[p + "n1.w"]), P(w[p + "n1.b"]), P(theta),
                     P(mem), P(memp), None, None, B, L, d, H, LN_EPS)
                tag("v", after=[p + "gate_vec"])
            else:
                call(p + "gate_scores", lib.svol_gate_scores, P(xp_cur), P(u[li]), P(scores), B, L, d, H)
                tag("v", after=[p + "gate_vec"])
                if self.plain:
                    call(p + "gate_apply", lib.svol_gate_apply, P(x_cur), P(scores), P(w[p + "n1.w"]), P(w[p + "n1.b"]), P(pos),
                         P(mem), P(memp), None, B, L, d, H, LN_EPS)
                else:
                    call(p + "gate_apply", lib.svol_gate_apply_theta, P(x_cur), P(scores), P(w[p + "n1.w"]), P(w[p + "n1.b"]),
                         P(theta), P(mem), P(memp), None, B, L, d, H, LN_EPS)
            # (b) video self-attention + norm2, FFN + norm3                               :137-143
            if self.plain:
                gemm(p + "sa_qk", memp, w[p + "sa.wqk"], w[p + "sa.bqk"], out=qk)
                gemm(p + "sa_v", mem, w[p + "sa.wv"], w[p + "sa.bv"], out_vt=vt, vt_len=L, vt_pitch=Lp)
            else:
                gemm(p + "sa_qkv", memp, w[p + "sa.wqkv"], w[p + "sa.bqkv"], out=qk, out_vt=vt, vt_len=L, vt_pitch=Lp,
                     A2=mem, split=2)
            attention(p + "sa_attn", qk, qk[:, d:], vt, att, L, L, 2 * d, 2 * d, Lp)
            gemm(p + "sa_out", att, w[p + "sa.wo"], w[p + "sa.bo"], out=mem2, residual=mem, ln=(w[p + "n2.w"], w[p + "n2.b"]))
            if self.plain:
                hid = plan.buf["hid"] if "hid" in plan.buf else buf("hid", (M, ff), bf)
                gemm(p + "ffn1_up", mem2, w[p + "mlp1.w1"], w[p + "mlp1.b1"], out=hid, act=ACT_GELU)
                gemm(p + "ffn1_down", hid, w[p + "mlp1.w2"], w[p + "mlp1.b2"], out=X, residual=mem2,
                     ln=(w[p + "n3.w"], w[p + "n3.b"]), out_pos=Xp, pos_t=pos)
            else:       # fused fc1 -> GELU -> fc2 -> +residual -> norm3, sine positions evaluated in the epilogue
                ffn(p + "ffn1", mem2, w[p + "mlp1.w1"], w[p + "mlp1.b1"], w[p + "mlp1.w2"], w[p + "mlp1.b2"],
                    (w[p + "n3.w"], w[p + "n3.b"]), out=X, out_pos=Xp, theta_t=theta)
            x_cur, xp_cur = X, Xp            # layer output mem (and mem + pos)
            # K / V^T of the cross-attention (:151-154) belong to the frame-token branch
            if self.plain:
                gemm(p + "ca_k", Xp, w[p + "ca.wk"], w[p + "ca.bk"], out=kc[li])
                gemm(p + "ca_v", X, w[p + "ca.wv"], w[p + "ca.bv"], out_vt=vt_ca[li], vt_len=L, vt_pitch=Lp)
                kv_done = p + "ca_v"
            else:
                gemm(p + "ca_kv", Xp, w[p + "ca.wkv"], w[p + "ca.bkv"], out=kc[li], out_vt=vt_ca[li], vt_len=L, vt_pitch=Lp,
                     A2=X, split=1)
                kv_done = p + "ca_kv"
            # ---- object-query branch
            q_first = len(plan.calls)
            # (c) query self-attention + norm4                                            :145-149
            if self.plain:
                gemm(p + "ta_qk", outp_cur, w[p + "ta.wqk"], w[p + "ta.bqk"], out=qkq)
                gemm(p + "ta_v", out_cur, w[p + "ta.wv"], w[p + "ta.bv"], out_vt=vtq, vt_len=Q, vt_pitch=Qp)
            else:
                gemm(p + "ta_qkv", outp_cur, w[p + "ta.wqkv"], w[p + "ta.bqkv"], out=qkq, out_vt=vtq, vt_len=Q, vt_pitch=Qp,
                     A2=out_cur, split=2)
            attention(p + "ta_attn", qkq, qkq[:, d:], vtq, attq, Q, Q, 2 * d, 2 * d, Qp)
            gemm(p + "ta_out", attq, w[p + "ta.wo"], w[p + "ta.bo"], out=o1, residual=out_cur,
                 ln=(w[p + "n4.w"], w[p + "n4.b"]), out_pos=o1p, pos_t=w["query_embed"], pos_mod=Q)
            # (d) query -> video cross-attention (padded keys masked) + norm5, FFN + norm6 :151-158
            gemm(p + "ca_q", o1p, w[p + "ca.wq"], w[p + "ca.bq"], out=qc)
            attention(p + "ca_attn", qc, kc[li], vt_ca[li], attq, Q, L, d, d, Lp, mask=vmask)
            plan.after[p + "ca_attn"] = [kv_done]
            gemm(p + "ca_out", attq, w[p + "ca.wo"], w[p + "ca.bo"], out=o2, residual=o1, ln=(w[p + "n5.w"], w[p + "n5.b"]))
            if self.plain:
                hidq = plan.buf["hidq"] if "hidq" in plan.buf else buf("hidq", (MQ, ff), bf)
                gemm(p + "ffn2_up", o2, w[p + "mlp2.w1"], w[p + "mlp2.b1"], out=hidq, act=ACT_GELU)
                gemm(p + "ffn2_down", hidq, w[p + "mlp2.w2"], w[p + "mlp2.b2"], out=hs[li], residual=o2,
                     ln=(w[p + "n6.w"], w[p + "n6.b"]), out_pos=outp, pos_t=w["query_embed"], pos_mod=Q)
            else:
                # (out + query_pos feeds the NEXT layer's query self-attention: not produced by the last layer)
                last_layer = li == NL - 1
                ffn(p + "ffn2", o2, w[p + "mlp2.w1"], w[p + "mlp2.b1"], w[p + "mlp2.w2"], w[p + "mlp2.b2"],
                    (w[p + "n6.w"], w[p + "n6.b"]), out=hs[li], out_pos=None if last_layer else outp,
                    pos_t=None if last_layer else w["query_embed"], pos_mod=0 if last_layer else Q)
            for name, _, _ in plan.calls[q_first:]:
                plan.branch[name] = "q"
            out_cur, outp_cur = hs[li], outp
        # ---- heads on every layer's queries (svanet.py:125-127)
        hs_all = hs.view(NL * MQ, d)
        gemm("box0", hs_all, w["box.0.w"], w["box.0.b"], out=h1, act=ACT_RELU)
        gemm("box1", h1, w["box.1.w"], w["box.1.b"], out=h2, act=ACT_RELU)
        call("heads", lib.svol_heads, P(hs_all), P(h2), P(w["cls.w"]), P(w["cls.b"]), P(w["box.2.w"]), P(w["box.2.b"]),
             P(logits), P(boxes), NL * MQ, d)
        for name in ("box0", "box1", "heads"):
            plan.branch[name] = "q"
        return plan

    # ------------------------------------------------------------------ run
    def plan_for(self, B: int, L: int, d_in: int) -> _Plan:
        """The plan (workspace + recorded launches + CUDA graph) of this input shape ON THE CURRENT STREAM: a caller that
        keeps two batches in flight on two streams gets two independent workspaces, so the low-occupancy tail of one
        forward (the object-query chain, the heads) overlaps the large frame-token kernels of the next."""
        self._weights()
        key = (B, L, d_in, _lib.stream_ptr())
        plan = self._plans.get(key)
        if plan is None:
            plan = self._build_plan(B, L, d_in)
            self._plans[key] = plan
        self.launches_per_forward = len(plan.calls)
        return plan

    @torch.no_grad()
    def forward(self, src_sketch, src_sketch_mask, src_video, src_video_mask):
        """Returns (logits [NL,B,Q,2], boxes [NL,B,Q,4]) fp32, views of the plan's static output buffers
        (overwritten by the next forward of the same shape)."""
        _lib.require_device()
        fmap = src_video.dim() == 5
        if fmap:
            # backbone hand-off: (B, T, C, h, w) trunk output, tokens = (frame, position) in the
            # reference's flatten(2).transpose(1, 2).reshape(N, -1, C) order (backbone.py:72-89)
            B, T, d_in, fh, fw = src_video.shape
            L = T * fh * fw
        else:
            B, L, d_in = src_video.shape
        plan = self.plan_for(B, L, d_in)
        b = plan.buf
        # The first kernel (LayerNorm of the frame tokens, the only reader of the 100 MB fp32 input) always runs
        # eagerly and reads the caller's tensor in place; everything after it works on plan-owned buffers and can
        # be replayed as one CUDA graph.
        src = src_video
        if (not fmap and self.use_graph and src.data_ptr() == b["src_video"].data_ptr()
                and src_sketch.data_ptr() == b["src_sketch"].data_ptr()
                and src_video_mask.data_ptr() == b["src_video_mask"].data_ptr()):
            # the caller filled the plan's own input buffers (input_buffers()): nothing to copy, and the input LayerNorm is
            # replayed as part of the graph -- one graph launch per forward
            self.run_plan(plan, full=True)
            return b["logits"], b["boxes"]
        bf16_in = (not fmap) and src.dtype == torch.bfloat16      # frame features kept in bf16 by the caller (feature cache)
        if fmap:
            if not (src.is_cuda and src.dtype == torch.float32 and src.is_contiguous()):
                src = src.to(device=b["src_video"].device, dtype=torch.float32).contiguous()
        elif bf16_in:
            if not (src.is_cuda and src.is_contiguous() and src.data_ptr() % 16 == 0):
                src = src.to(device=b["src_video"].device).contiguous()
        elif not (src.is_cuda and src.dtype == torch.float32 and src.is_contiguous() and src.data_ptr() % 16 == 0):
            b["src_video"].copy_(src_video, non_blocking=True)
            src = b["src_video"]
        if src_sketch.data_ptr() != b["src_sketch"].data_ptr():
            if src_sketch.dim() == 3 and src_sketch.shape[1] != 1:
                raise NotImplementedError("svol_b200 supports one sketch token per pair (L_sketch == 1)")
            b["src_sketch"].copy_(src_sketch.reshape(B, -1), non_blocking=True)
        if src_video_mask.data_ptr() != b["src_video_mask"].data_ptr():
            b["src_video_mask"].copy_(src_video_mask, non_blocking=True)
        assert plan.ln_in_index == 0
        name, fn, args = plan.calls[0]
        if fmap:   # LayerNorm straight from the channel-major feature map: no permuted fp32 copy
            rc = _lib.get_lib().svol_layernorm_nchw_to_bf16(src.data_ptr(), args[1], args[2], args[3], B * T, d_in, fh * fw, LN_EPS,
                                                            _lib.stream_ptr())
        elif bf16_in:
            rc = _lib.get_lib().svol_layernorm_bf16_to_bf16(src.data_ptr(), args[1], args[2], args[3], B * L, d_in, LN_EPS,
                                                            _lib.stream_ptr())
        else:
            rc = fn(src.data_ptr(), *args[1:], _lib.stream_ptr())
        if rc != 0:
            _lib.check(rc, name)
        self.run_plan(plan)
        return b["logits"], b["boxes"]

    def input_buffers(self, B: int, L: int, d_in: int) -> Dict[str, torch.Tensor]:
        """The static input tensors of this shape's plan ON THE CURRENT STREAM: ``src_video`` (B,L,d_in) fp32,
        ``src_sketch`` (B,d_sk) fp32 (pass it to forward as ``.view(B, 1, -1)`` or as is), ``src_video_mask`` (B,L) fp32.
        A caller that produces its inputs on the device (a backbone, a feature cache) writes them here and passes these
        very tensors to ``forward``: the forward is then a single CUDA-graph launch (input LayerNorm included)."""
        b = self.plan_for(B, L, d_in).buf
        return {k: b[k] for k in ("src_video", "src_sketch", "src_video_mask")}

    def run_plan(self, plan: _Plan, full: bool = False) -> None:
        """Runs calls[1:] of the plan (everything after the input LayerNorm; ``full``: calls[0:], the LayerNorm reading the
        plan's own input buffer), eagerly or as a graph replay."""
        if self.use_graph:
            attr, start = ("graph_full", 0) if full else ("graph", 1)
            if getattr(plan, attr) is None:
                # warm-up run outside capture (sets function attributes, loads modules)
                plan.run(_lib.stream_ptr(), start=start)
                torch.cuda.synchronize()
                if self._side is None:
                    # the query / input branches are captured on high-priority streams (SVOL_B200_SIDE_PRIORITY, default -1; their kernel
                    # nodes inherit it), so that the short query-side kernels get SMs ahead of the frame-token kernels
                    prio = int(os.environ.get("SVOL_B200_SIDE_PRIORITY", "-1"))
                    self._side = (torch.cuda.Stream(priority=prio), torch.cuda.Stream(priority=prio))
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    plan.run(_lib.stream_ptr(), start=start, side=self._side)
                setattr(plan, attr, g)
            getattr(plan, attr).replay()
        else:
            plan.run(_lib.stream_ptr(), start=1)
