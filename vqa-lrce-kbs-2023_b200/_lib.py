"""ctypes binding of liblrce_b200.so (the C ABI declared in include/lrce_b200.h).

There is deliberately no fallback: if the shared library has not been built (``python -c "import __graft_entry__ as g;
g.build()"`` or ``make -C vqa-lrce-kbs-2023_b200/csrc``) loading raises, and every entry point raises LrceError on a
non-zero return code with the library's own message.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# LRCE_B200_LIB: another build of the same library (timing variants of tools/build_variants.sh); it must exist, there is no fallback
LIB_PATH = os.environ.get("LRCE_B200_LIB") or os.path.join(_HERE, "liblrce_b200.so")

_c = ctypes
_vp, _i, _f, _ll, _u64, _sz = _c.c_void_p, _c.c_int, _c.c_float, _c.c_longlong, _c.c_ulonglong, _c.c_size_t

# name -> argtypes (every function returns int unless listed in _RESTYPES)
_SIGNATURES = {
    "lrce_abi_version": [],
    "lrce_last_error": [],
    "lrce_gemm_bf16": [_vp, _i, _vp, _i, _i, _i, _i, _vp, _vp, _i, _vp, _i, _i, _i, _vp, _vp, _f, _vp, _i, _vp, _f, _vp, _vp],
    "lrce_mlp_fused_bf16": [_vp, _i, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _i, _vp, _i, _i, _vp],
    "lrce_mlp_l2_scratch_bytes": [_i],
    "lrce_mlp_l2_bf16": [_vp, _i, _vp, _vp, _vp, _vp, _i, _f, _vp, _vp, _vp, _i, _vp, _vp, _sz, _i, _i, _vp],
    "lrce_layernorm_bf16": [_vp, _vp, _vp, _vp, _f, _ll, _i, _i, _vp],
    "lrce_patch_merge_ln_bf16": [_vp, _vp, _vp, _vp, _f, _i, _i, _i, _i, _i, _vp],
    "lrce_patch_gather_f32": [_vp, _vp, _i, _i, _i, _i, _vp],
    "lrce_patch_gather_u8": [_vp, _vp, _i, _i, _i, _i, _vp],
    "lrce_window_remap_bf16": [_vp, _vp] + [_i] * 12 + [_vp],
    "lrce_remap_index": [_vp, _vp, _vp] + [_i] * 9 + [_vp],
    "lrce_window_bias_pack": [_vp, _vp, _i, _vp],
    "lrce_window_attention_bf16": [_vp, _vp, _vp] + [_i] * 8 + [_vp],
    "lrce_video_posembed_ln": [_vp] * 7 + [_f, _vp, _i, _i, _i, _i, _vp],
    "lrce_text_posembed_ln": [_vp, _i] + [_vp] * 4 + [_f, _vp, _i, _i, _vp],
    "lrce_encoder_walk_pack_bytes": [_i, _i],
    "lrce_encoder_walk_pack": [_vp, _i, _vp, _vp, _i, _vp, _vp],
    "lrce_window_attention_profile": [_vp, _vp, _vp] + [_i] * 8 + [_vp, _vp],
    "lrce_encoder_walk": [_vp, _i, _vp, _vp, _i, _vp, _vp, _vp, _f, _i, _i, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "lrce_encoder_walk_plan": [_i, _i, _vp, _vp, _vp, _vp, _vp],
    "lrce_encoder_walk_profile": [_vp, _i, _vp, _vp, _i, _vp, _vp, _vp, _f, _i, _i, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _i],
    "lrce_add_ln_768": [_vp, _vp, _i, _vp, _vp, _f, _vp, _vp, _vp, _vp, _i, _i, _ll, _f, _i, _f, _i, _u64, _vp, _vp],
    "lrce_ln_bwd_768": [_vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _i, _i, _vp, _vp, _ll, _f, _i, _f, _i, _u64, _vp, _vp],
    "lrce_rows_f32_to_bf16": [_vp, _vp, _vp, _vp, _i, _i, _ll, _i, _i, _i, _f, _i, _u64, _vp, _vp],
    "lrce_dropout_bf16": [_vp, _ll, _f, _i, _u64, _vp, _vp],
    "lrce_xattn_fwd": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _f, _i, _u64, _vp, _vp],
    "lrce_xattn_bwd": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _f, _i, _u64, _vp, _vp],
    "lrce_rowsum_bf16": [_vp, _i, _i, _vp, _i, _i, _vp],
    "lrce_colsum": [_vp, _i, _ll, _i, _ll, _vp, _vp],
    "lrce_transpose_bf16": [_vp, _ll, _i, _ll, _vp, _ll, _vp],
    "lrce_add_bf16": [_vp, _vp, _vp, _vp, _ll, _vp],
    "lrce_posembed_bwd": [_vp, _vp, _vp, _i] + [_vp] * 5 + [_f] + [_vp] * 7 + [_i] * 5 + [_f, _i, _u64, _vp, _vp],
    "lrce_gemm_skinny_bf16": [_vp, _i, _vp, _i, _i, _i, _i, _vp, _vp, _i, _vp],
    "lrce_bert_embed_ln": [_vp] * 7 + [_f, _vp, _vp, _ll, _i, _i, _i, _i, _vp],
    "lrce_bert_attention": [_vp, _vp, _vp, _i, _i, _i, _vp],
}
_RESTYPES = {"lrce_last_error": _c.c_char_p, "lrce_encoder_walk_pack_bytes": _c.c_size_t, "lrce_mlp_l2_scratch_bytes": _c.c_size_t}

_lib = None


class LrceError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LrceError(
                f"{LIB_PATH} is missing: build the CUDA extension first (python -c 'import __graft_entry__ as g; "
                "g.build()'). There is no CPU or PyTorch fallback for the LRCE hot path.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, argtypes in _SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is not exported
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, _c.c_int)
        _lib = handle
    return _lib


def exported_symbols():
    return sorted(_SIGNATURES)


def check(rc, what):
    if rc != 0:
        msg = lib().lrce_last_error()
        raise LrceError(f"{what} failed with code {rc}: {msg.decode() if msg else '?'}")
