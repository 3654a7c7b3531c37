"""ctypes binding of libvcoder_b200.so — the binding a reference maintainer would add.

The library is built in-tree by `vcoder_b200.build`; importing this module never builds and never falls back to
anything else: if the .so is missing, loading raises.  Symbol signatures mirror include/vcoder_b200.h one to one.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "libvcoder_b200.so"

VCB_OK = 0
ERR_NAMES = {
    -1: "VCB_ERR_INVALID", -2: "VCB_ERR_CUDA", -3: "VCB_ERR_STATE", -4: "VCB_ERR_NOMEM", -5: "VCB_ERR_RAGGED",
    -6: "VCB_ERR_SPECIAL_INDEX", -7: "VCB_ERR_NO_DEVICE",
}


class VcbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class VcbConfig(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("variant", C.c_int32),
        ("vit_image_size", C.c_int32), ("vit_patch", C.c_int32), ("vit_hidden", C.c_int32), ("vit_heads", C.c_int32),
        ("vit_ffn", C.c_int32), ("vit_layers", C.c_int32), ("vit_select_layer", C.c_int32), ("vit_ln_eps", C.c_float),
        ("llm_hidden", C.c_int32), ("llm_layers", C.c_int32), ("llm_heads", C.c_int32), ("llm_ffn", C.c_int32),
        ("llm_vocab", C.c_int32), ("llm_rms_eps", C.c_float), ("llm_rope_theta", C.c_float),
        ("max_batch", C.c_int32), ("max_seq_len", C.c_int32), ("kv_page_tokens", C.c_int32),
    ]


class VcbTensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("ndim", C.c_int32), ("shape", C.c_int64 * 4)]


_p = C.c_void_p
_i32 = C.c_int32
_i64 = C.c_int64
_f = C.c_float

# int cb(void* user, int32 step, const int32* ids_host, int32 B)  (vcb_token_callback)
TOKEN_CALLBACK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.c_int32)

# name -> (restype, argtypes): every symbol declared in include/vcoder_b200.h
SIGNATURES = {
    "vcb_version": (C.c_char_p, []),
    "vcb_last_error": (C.c_char_p, []),
    "vcb_device_count": (C.c_int, []),
    "vcb_op_gemm_f16": (C.c_int, [_p, _i32, _p, _i32, _i32, _i32, _i32, _p, _p, _i32, _i32, _p, _i32, _i32, _p]),
    "vcb_op_skinny_gemm_f16": (C.c_int, [_p, _i32, _p, _i32, _i32, _i32, _p, _i32, _i32, _p, _i32, _p]),
    "vcb_op_skinny_gemm_onetile_f16": (C.c_int, [_p, _i32, _p, _i32, _i32, _i32, _p, _i32, _i32, _p, _i32, _p]),
    "vcb_op_pack_skinny_weights": (C.c_int, [_p, _i32, _i32, _p, _p]),
    "vcb_op_skinny_gemm_packed_f16": (C.c_int, [_p, _i32, _p, _i32, _i32, _i32, _p, _i32, _i32, _p, _i32, _p]),
    "vcb_op_skinny_rmsnorm_gemm_f16": (C.c_int, [_p, _i32, _p, _f, _p, _i32, _i32, _i32, _p, _i32, _i32, _p, _i32, _p]),
    "vcb_op_layernorm_f16": (C.c_int, [_p, _p, _p, _p, _i64, _i32, _f, _p]),
    "vcb_op_rmsnorm_f16": (C.c_int, [_p, _p, _p, _i64, _i32, _f, _p]),
    "vcb_op_attention_vit_f16": (C.c_int, [_p, _i32, _i32, _i32, _f, _p, _p]),
    "vcb_op_attention_prefill_f16": (C.c_int, [_p, _i64, _p, _p, _p, _i32, _i32, _i32, _i32, _i32, _f, _p, _i64, _p]),
    "vcb_op_attention_decode_f16": (C.c_int, [_p, _p, _p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _f, _f, _p, _p]),
    "vcb_op_nf4_roundtrip_f16": (C.c_int, [_p, _i64, _i32, _p, _p, _p, _p]),
    "vcb_splice_plan": (C.c_int, [_i32, _p, _i32, _i32, _i32, _i32, _p, _p, _i64, _p]),
    "vcb_create": (C.c_int, [C.POINTER(VcbConfig), _i32, C.POINTER(_p)]),
    "vcb_destroy": (C.c_int, [_p]),
    "vcb_bind_weights": (C.c_int, [_p, C.POINTER(VcbTensor), _i32]),
    "vcb_tower_forward": (C.c_int, [_p, _p, _i32, _p, _p]),
    "vcb_encode": (C.c_int, [_p, _i32, _p, _i32, _p, _p]),
    "vcb_prefill": (C.c_int, [_p, _p, _i32, _i32, _p, _p, _p, _p, _p, _p, _p, _p]),
    "vcb_forward_logits": (C.c_int, [_p, _p, _i32, _i32, _p, _p, _p, _p, C.c_int64, _p, _p]),
    "vcb_splice_embeds": (C.c_int, [_p, _p, _i32, _i32, _p, _p, _p, _p, _p, _p]),
    "vcb_prefill_cached": (C.c_int, [_p, _p, _i32, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "vcb_prefix_match": (C.c_int, [_p, _i32, _p, _i32, _i32, _i32, _i32, _p]),
    "vcb_decode": (C.c_int, [_p, _i32, _p, _p, _p]),
    "vcb_generate": (C.c_int, [_p, _p, _i32, _i32, _p, _p, _p, _i32, _p, _p]),
    "vcb_set_sampling": (C.c_int, [_p, _i32, _f, _i32, _f, C.c_uint64, _p]),
    "vcb_decode_stream": (C.c_int, [_p, _i32, _i32, TOKEN_CALLBACK, _p, _p, _p]),
    "vcb_op_sample_f16": (C.c_int, [_p, _i32, _i32, _f, _i32, _f, C.c_uint64, _i32, _p, _p, _p]),
    "vcb_op_preprocess_u8": (C.c_int, [_p, _i32, _i32, _p, _i32, _p, _p, _p, _p]),
    "vcb_philox_uniform": (C.c_float, [C.c_uint64, C.c_uint32, C.c_uint32]),
    "vcb_microbench": (C.c_int, [_i32, _i32, _i32, _i32, _p]),
    "vcb_tune_skinny": (C.c_int, [_i32, _i32, _i32, _i32]),
    "vcb_trace_begin": (C.c_int, [_i32]),
    "vcb_trace_read": (C.c_int, [_p, _i32, _p, _p, _i64]),
    "vcb_bench_weight_stream": (C.c_int, [_p, _i32, _i32, _p, _p, _p]),
    "vcb_reset": (C.c_int, [_p]),
    "vcb_launch_count": (C.c_uint64, []),
    "vcb_profile_enable": (C.c_int, [_i32]),
    "vcb_profile_read": (C.c_int, [_p, _p, _i32]),
}

PROFILE_CATEGORIES = ["gemm_tc5", "skinny_gemm", "attn_vit", "attn_prefill", "attn_decode", "norm", "rope_kv", "misc"]

_lib = None


def load() -> C.CDLL:
    """dlopen the in-tree library (no build, no fallback)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise FileNotFoundError(
                f"{LIB_PATH} is missing: build it with `python -m vcoder_b200.build` (there is no fallback path)")
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != VCB_OK:
        raise VcbError(rc, load().vcb_last_error().decode("utf-8", "replace"))


def ptr(t) -> C.c_void_p:
    """data pointer of a torch tensor / numpy array / None."""
    if t is None:
        return C.c_void_p(0)
    if hasattr(t, "data_ptr"):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)
