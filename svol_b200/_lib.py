"""ctypes binding of ``libsvol_b200.so`` (the C ABI declared in ``include/svol_b200.h``).

There is no fallback: if the shared object is missing or the device is not a B200-class GPU,
every entry point raises.  The structures below mirror the header field for field.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libsvol_b200.so")

ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
ABI_VERSION = 13

# every symbol include/svol_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "svol_abi_version", "svol_last_error", "svol_device_check", "svol_sizeof_args",
    "svol_gemm_bf16", "svol_gemm_bf16_plain", "svol_ffn_bf16", "svol_attention_bf16", "svol_attention_bf16_plain",
    "svol_layernorm_f32_to_bf16", "svol_ln_linear_f32", "svol_posenc_sine", "svol_posenc_theta", "svol_add_pos_bf16",
    "svol_gate_vectors", "svol_gate_scores", "svol_gate_apply", "svol_gate_apply_theta", "svol_gate_fused",
    "svol_gate_fused_supported", "svol_heads",
    "svol_match", "svol_match_localize", "svol_lsap_f32", "svol_criterion", "svol_criterion_backward", "svol_criterion_scratch_bytes", "svol_postprocess",
    # training step
    "svol_layernorm_bf16", "svol_layernorm_backward", "svol_gelu_bf16", "svol_act_backward", "svol_transpose_bf16",
    "svol_colsum_bf16", "svol_attention_backward_bf16", "svol_heads_backward", "svol_gate_backward",
    "svol_gate_vectors_backward", "svol_ln_linear_f32_backward", "svol_batch_sum", "svol_accum_bf16", "svol_adamw", "svol_adamw_segments",
    "svol_pack_weights", "svol_layernorm_f32_to_bf16_dropout", "svol_ln_linear_f32_dropout", "svol_layernorm_nchw_to_bf16", "svol_layernorm_bf16_to_bf16",
    "svol_eval_max_iou", "svol_eval_average_precision",
]


class GemmEpilogue(C.Structure):
    _fields_ = [
        ("bias", C.c_void_p), ("act", C.c_int32), ("ld_res", C.c_int32), ("residual", C.c_void_p),
        ("ln_weight", C.c_void_p), ("ln_bias", C.c_void_p), ("ln_eps", C.c_float), ("ld_out", C.c_int32),
        ("out", C.c_void_p), ("out_pos", C.c_void_p), ("pos", C.c_void_p), ("ld_pos", C.c_int32),
        ("pos_row_mod", C.c_int32), ("out_vt", C.c_void_p), ("vt_len", C.c_int32), ("vt_pitch", C.c_int32),
        ("pos_theta", C.c_void_p),
        ("out_pre", C.c_void_p), ("dact_src", C.c_void_p), ("ld_dact", C.c_int32), ("dact_mode", C.c_int32),
    ]


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("W", C.c_void_p), ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("lda", C.c_int32), ("ldw", C.c_int32), ("split_block", C.c_int32), ("ep", GemmEpilogue),
        ("A2", C.c_void_p), ("lda2", C.c_int32), ("ld_f32", C.c_int32), ("out_f32", C.c_void_p),
        ("mn_major", C.c_int32), ("reserved", C.c_int32),
    ]


class FfnArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p),
        ("ln_weight", C.c_void_p), ("ln_bias", C.c_void_p), ("out", C.c_void_p), ("out_pos", C.c_void_p),
        ("pos", C.c_void_p), ("pos_theta", C.c_void_p),
        ("M", C.c_int32), ("d", C.c_int32), ("ff", C.c_int32), ("ldx", C.c_int32), ("ldw1", C.c_int32),
        ("ldw2", C.c_int32), ("ld_out", C.c_int32), ("ld_pos", C.c_int32), ("pos_row_mod", C.c_int32),
        ("ln_eps", C.c_float), ("reserved", C.c_int32),
    ]


class AttnArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("vt", C.c_void_p), ("key_mask", C.c_void_p), ("out", C.c_void_p),
        ("B", C.c_int32), ("H", C.c_int32), ("Lq", C.c_int32), ("Lk", C.c_int32), ("ldq", C.c_int32),
        ("ldk", C.c_int32), ("ldo", C.c_int32), ("vt_pitch", C.c_int32),
        ("lse", C.c_void_p), ("lse_pitch", C.c_int32), ("reserved", C.c_int32),
    ]


class AttnBwdArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("kt", C.c_void_p), ("qt", C.c_void_p),
        ("o", C.c_void_p), ("d_o", C.c_void_p), ("d_ot", C.c_void_p), ("lse", C.c_void_p), ("delta", C.c_void_p),
        ("key_mask", C.c_void_p), ("dq", C.c_void_p), ("dk", C.c_void_p), ("dv", C.c_void_p),
        ("B", C.c_int32), ("H", C.c_int32), ("Lq", C.c_int32), ("Lk", C.c_int32), ("ldq", C.c_int32),
        ("ldk", C.c_int32), ("ldv", C.c_int32), ("ld_o", C.c_int32), ("ld_do", C.c_int32), ("ld_dq", C.c_int32),
        ("ld_dk", C.c_int32), ("ld_dv", C.c_int32), ("kt_pitch", C.c_int32), ("qt_pitch", C.c_int32),
        ("stat_pitch", C.c_int32), ("reserved", C.c_int32),
    ]


class PackJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("rows", C.c_int32), ("cols", C.c_int32),
                ("scaled_rows", C.c_int32), ("flags", C.c_int32), ("scale", C.c_float), ("reserved", C.c_int32)]


PACK_BF16, PACK_TRANSPOSE = 1, 2


class AdamwGroup(C.Structure):
    _fields_ = [("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("weight_decay", C.c_float), ("step", C.c_int32)]


class MatchArgs(C.Structure):
    _fields_ = [
        ("logits", C.c_void_p), ("boxes", C.c_void_p), ("tgt_boxes", C.c_void_p), ("tgt_off", C.c_void_p),
        ("match_off", C.c_void_p), ("cost_off", C.c_void_p), ("cost_ws", C.c_void_p), ("pred_idx", C.c_void_p),
        ("tgt_idx", C.c_void_p), ("status", C.c_void_p),
        ("NL", C.c_int32), ("B", C.c_int32), ("Q", C.c_int32), ("problems_per_video", C.c_int32),
        ("rows_per_problem", C.c_int32), ("max_cols", C.c_int32),
        ("w_class", C.c_float), ("w_bbox", C.c_float), ("w_giou", C.c_float), ("K", C.c_int32),
        ("video_match_off", C.c_void_p), ("video_tgt_off", C.c_void_p),
        ("mode", C.c_int32), ("solver", C.c_int32), ("localize", C.c_int32), ("reserved", C.c_int32),
        ("order", C.c_void_p),
    ]


class CriterionArgs(C.Structure):
    _fields_ = [
        ("logits", C.c_void_p), ("boxes", C.c_void_p), ("tgt_boxes", C.c_void_p), ("pred_idx", C.c_void_p),
        ("tgt_idx", C.c_void_p), ("match_video", C.c_void_p), ("video_tgt_off", C.c_void_p), ("losses", C.c_void_p),
        ("NL", C.c_int32), ("B", C.c_int32), ("Q", C.c_int32), ("K", C.c_int32),
        ("eos_coef", C.c_float), ("idx_pitch", C.c_int32), ("video_match_off", C.c_void_p), ("meta", C.c_void_p),
        ("scratch", C.c_void_p),
    ]


_lib: Optional[C.CDLL] = None
_i32, _i64, _f32, _vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p


def _declare(lib: C.CDLL) -> None:
    lib.svol_abi_version.restype = C.c_int
    lib.svol_last_error.restype = C.c_char_p
    lib.svol_device_check.restype = C.c_int
    sigs = {
        "svol_gemm_bf16": [C.POINTER(GemmArgs), _vp],
        "svol_gemm_bf16_plain": [C.POINTER(GemmArgs), _vp],
        "svol_ffn_bf16": [C.POINTER(FfnArgs), _vp],
        "svol_attention_bf16": [C.POINTER(AttnArgs), _vp],
        "svol_attention_bf16_plain": [C.POINTER(AttnArgs), _vp],
        "svol_layernorm_f32_to_bf16": [_vp, _vp, _vp, _vp, _i32, _i32, _f32, _vp],
        "svol_ln_linear_f32": [_vp, _vp, _vp, _vp, _vp, _i32, _vp, _i32, _i32, _i32, _f32, _vp],
        "svol_layernorm_nchw_to_bf16": [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _f32, _vp],
        "svol_layernorm_bf16_to_bf16": [_vp, _vp, _vp, _vp, _i32, _i32, _f32, _vp],
        "svol_posenc_sine": [_vp, _vp, _i32, _i32, _i32, _vp],
        "svol_posenc_theta": [_vp, _vp, _i32, _i32, _vp],
        "svol_add_pos_bf16": [_vp, _vp, _vp, _i32, _i32, _i32, _vp],
        "svol_gate_vectors": [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp],
        "svol_gate_scores": [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp],
        "svol_gate_apply": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _vp],
        "svol_gate_apply_theta": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _vp],
        "svol_gate_fused": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _vp],
        "svol_gate_fused_supported": [_i32],
        "svol_heads": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp],
        "svol_match": [C.POINTER(MatchArgs), _vp],
        "svol_match_localize": [_vp, _vp, _i32, _i32, _i32, _vp],
        "svol_lsap_f32": [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _vp],
        "svol_criterion": [C.POINTER(CriterionArgs), _vp],
        "svol_criterion_backward": [C.POINTER(CriterionArgs), _vp, _vp, _vp, _vp],
        "svol_criterion_scratch_bytes": [_i32, _i32],
        "svol_postprocess": [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp],
        "svol_layernorm_bf16": [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _i32, _i32, _f32, _f32, _vp, _i32, _vp],
        "svol_layernorm_backward": [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _f32, _f32, _vp, _i32, _vp],
        "svol_layernorm_f32_to_bf16_dropout": [_vp, _vp, _vp, _vp, _i32, _i32, _f32, _f32, _vp, _i32, _vp],
        "svol_ln_linear_f32_dropout": [_vp, _vp, _vp, _vp, _vp, _i32, _vp, _i32, _i32, _i32, _f32, _f32, _vp, _i32, _vp],
        "svol_gelu_bf16": [_vp, _vp, _i64, _vp],
        "svol_act_backward": [_vp, _vp, _vp, _i64, _i32, _vp],
        "svol_transpose_bf16": [_vp, _i32, _i32, _i32, _vp, _i32, _vp, _vp],
        "svol_colsum_bf16": [_vp, _i32, _i32, _i32, _vp, _vp],
        "svol_attention_backward_bf16": [C.POINTER(AttnBwdArgs), _vp],
        "svol_heads_backward": [_vp] * 13 + [_i32, _i32, _vp],
        "svol_gate_backward": [_vp] * 8 + [_i32, _i32, _i32, _i32, _vp],
        "svol_gate_vectors_backward": [_vp] * 7 + [_i32, _i32, _i32, _vp],
        "svol_ln_linear_f32_backward": [_vp] * 6 + [_i32] + [_vp] * 5 + [_i32, _i32, _i32, _f32, _f32, _vp, _i32, _vp],
        "svol_batch_sum": [_vp, _vp, _i32, _i32, _i32, _vp],
        "svol_accum_bf16": [_vp, _vp, _i64, _f32, _i32, _vp],
        "svol_adamw": [_vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _f32, _i32, _f32, _vp],
        "svol_adamw_segments": [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _i32, _vp, _i32, _f32, _vp],
        "svol_pack_weights": [_vp, _i32, _vp],
        "svol_eval_max_iou": [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp],
        "svol_eval_average_precision": [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp],
    }
    for name, argtypes in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    lib.svol_criterion_scratch_bytes.restype = C.c_int64


def get_lib() -> C.CDLL:
    """Loads the shared object (once).  Raises if it has not been built -- there is no CPU path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with svol_b200/csrc/build.sh "
                "(or __graft_entry__.build()); svol_b200 has no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        _declare(lib)
        if lib.svol_abi_version() != ABI_VERSION:
            raise ImportError("libsvol_b200.so ABI version mismatch; rebuild it")
        for which, struct in enumerate((GemmArgs, AttnArgs, MatchArgs, CriterionArgs, GemmEpilogue, FfnArgs, AttnBwdArgs, PackJob)):
            if lib.svol_sizeof_args(which) != C.sizeof(struct):
                raise ImportError(f"ctypes layout of {struct.__name__} does not match libsvol_b200.so")
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = get_lib().svol_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"svol_b200 {what} failed (code {rc}): {msg}")


_device_ok = set()


def require_device() -> None:
    """Fails loudly unless the current CUDA device can run the sm_100a kernels (checked once per device)."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("svol_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    dev = torch._C._cuda_getDevice()
    if dev not in _device_ok:
        check(get_lib().svol_device_check(), "device check")
        _device_ok.add(dev)


def ptr(t) -> Optional[int]:
    return None if t is None else t.data_ptr()


def stream_ptr() -> int:
    """cudaStream_t of torch's current stream on the current device (the raw C call: torch.cuda.current_stream() builds
    a Python Stream object and resolves the device index through several layers, ~8 us per call)."""
    import torch
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())
