"""ctypes binding of ``libvidi_b200.so`` (the C ABI in ``include/vidi_b200.h``).

The library is built in-tree (``vidi_b200/libvidi_b200.so``) by ``make -C vidi_b200/csrc`` /
``__graft_entry__.build()``.  There is no fallback: if the library is missing, or a call returns
non-zero, a ``RuntimeError`` is raised -- the product path never computes on the CPU or through
another library.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvidi_b200.so")

_p, _i, _l, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> argtypes, in the order of include/vidi_b200.h
SIGNATURES = {
    "vidi_gemm_bf16": [_p, _l, _p, _l, _p, _l, _i, _i, _i, _p, _p, _l, _i, _i, _f, _i, _i, _i, _p],
    "vidi_gemm_bf16_2cta": [_p, _l, _p, _l, _p, _l, _i, _i, _i, _p, _p, _l, _i, _i, _f, _i, _i, _i, _p],
    "vidi_gemm_bf16_2cta_ln": [_p, _l, _p, _l, _p, _l, _i, _i, _i, _p, _p, _l, _i, _i, _f, _i, _p, _i, _p, _f, _p, _p],
    "vidi_rmsnorm": [_p, _l, _p, _p, _l, _i, _i, _f, _i, _f, _p],
    "vidi_residual_norm": [_p, _l, _p, _l, _p, _p, _p, _l, _i, _i, _f, _i, _i, _p],
    "vidi_layernorm": [_p, _l, _p, _p, _p, _l, _i, _i, _f, _p],
    "vidi_mm_finish": [_p, _l, _p, _p, C.POINTER(_p), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), _i, _i, _i, _f,
                       _p, _l, _p, _i, _i, _f, _p],
    "vidi_rmsnorm_f32": [_p, _p, _i, _i, _f, _i, _p],
    "vidi_patch_im2col": [_p, _p, _i, _i, _i, _i, _p],
    "vidi_whisper_im2col1": [_p, _p, _i, _i, _i, _p],
    "vidi_whisper_im2col2": [_p, _p, _i, _i, _i, _p],
    "vidi_pool_s2d": [_p, _p, _i, _i, _i, _i, _i, _i, _p],
    "vidi_conv_window_gather": [_p, _p, _i, _i, _i, _i, _p],
    "vidi_bilinear_ac": [_p, _p, _i, _i, _i, _i, _p],
    "vidi_embed_gather": [_p, _p, _p, _i, _i, _i, _f, _p],
    "vidi_sinusoid_split": [_p, _p, _i, _i, _i, _i, _i, _p],
    "vidi_split3": [_p, _p, _l, _i, _i, _p],
    "vidi_cast_f32_bf16": [_p, _p, _l, _p],
    "vidi_resample_u8": [_p, _p, _l, _i, _i, _i, _p, _p, _i, _p],
    "vidi_resample_u8_to_chw_bf16": [_p, _p, _i, _i, _i, _i, _p, _p, _i, _f, _f, _f, _p],
    "vidi_logmel_frames": [_p, _p, _p, _i, _i, _p],
    "vidi_logmel_power": [_p, _l, _p, _l, _p],
    "vidi_logmel_finish": [_p, _i, _i, _p, _p, _p],
    "vidi_attn_dense": [_p, _l, _i, _i, _i, _p, _l, _i, _i, _i, _i, _f, _p],
    "vidi_attn_dense_v1": [_p, _l, _i, _i, _i, _p, _l, _i, _i, _i, _i, _f, _p],
    "vidi_attn_dense_v2": [_p, _l, _i, _i, _i, _p, _l, _i, _i, _i, _i, _f, _p],
    "vidi_attn_dense_poly": [_p, _l, _p, _l, _i, _i, _i, _i, _f, _i, _p],
    "vidi_attn_dense_mma": [_p, _l, _i, _i, _i, _p, _l, _i, _i, _i, _i, _f, _p],
    "vidi_xattn_splitkv": [_p, _l, _p, _p, _l, _p, _i, _i, _i, _i, _i, _i, _f, _f, _p, _p, _p],
    "vidi_xattn_splitkv_seg": [_p, _l, _p, _p, _l, _i, _i, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(_p),
                               _i, _i, _i, _i, _f, _f, _p, _p, _p],
    "vidi_xattn_splitkv_mma": [_p, _l, _p, _p, _l, _p, _i, _i, _i, _i, _i, _i, _f, _f, _p, _p, _p],
    "vidi_xattn_merge": [_p, _p, _i, _i, _l, _l, _i, _i, _f, _i, _p, _p],
    "vidi_text_qk_prep": [_p, _l, _p, _l, _p, _l, _i, _i, _i, _i, _p, _i, _p],
    "vidi_xattn_merge2": [_p, _p, _i, _i, _l, _l, _f, _p, _p, _i, _i, _l, _l, _f, _i, _p, _i, _i, _p, _p],
    "vidi_xattn_merge2_sync": [_p, _p, _i, _i, _l, _l, _f, _p, _p, _i, _i, _l, _l, _f, _i, _p, _i, _i, _p, _p, _i, C.c_uint32, _p, _p],
    "vidi_xattn_premerge_push": [_p, _p, _i, _p, _p, _i, _i, _i, _i, C.POINTER(_p), C.POINTER(_p), _i, _l, C.c_uint32, _p, _p],
    "vidi_p2p_alloc": [_l, C.POINTER(_p), _p],
    "vidi_p2p_open": [_p, C.POINTER(_p)],
    "vidi_p2p_close": [_p],
    "vidi_p2p_free": [_p],
    "vidi_rope_inplace": [_p, _l, _i, _i, _i, _i, _p, _i, _p],
    "vidi_attn_text": [_p, _l, _p, _p, _l, _i, _i, _i, _i, _i, _i, _f, _f, _i, _p, _p],
}


class VidiTextLayerW(C.Structure):
    _fields_ = [(n, _p) for n in ("wqkv", "wo", "wgu", "wd", "n_in", "n_post", "n_preff", "n_postff")]


class VidiTextSeg(C.Structure):
    _fields_ = [("row0", _l), ("rows", C.c_int32), ("splits", C.c_int32), ("kmask", _p), ("gate", _f), ("reserved", C.c_int32)]


class VidiTextPass(C.Structure):
    """mirror of ``struct VidiTextPass`` (include/vidi_b200.h); tests/test_host_cpu.py checks the layout against the C compiler"""
    _fields_ = ([(n, C.c_int32) for n in ("Tq", "pos0", "layers", "hidden", "heads", "kv_heads", "head_dim", "inter", "vocab", "gemma",
                                          "glu", "sliding_window", "logits_keep")]
                + [(n, _f) for n in ("rms_eps", "scale", "attn_softcap", "final_softcap", "normalizer")]
                + [("layer_w", C.POINTER(VidiTextLayerW)), ("embed", _p), ("final_norm", _p), ("lm_head", _p), ("inv_freq", _p), ("ids", _p),
                   ("text_kv", _p), ("text_kv_layer_stride", _l), ("text_kv_ld", _l),
                   ("stream_kv", _p), ("stream_layer_stride", _l), ("stream_ld", _l),
                   ("nseg", C.c_int32), ("stream_rows", C.c_int32), ("world", C.c_int32), ("rank", C.c_int32), ("seq0", C.c_uint32),
                   ("reserved0", C.c_int32),
                   ("seg", VidiTextSeg * 2), ("peer_data", _p * 16), ("peer_flags", _p * 16), ("cap", _l),
                   ("counter", _p), ("err", _p), ("workspace", _p), ("workspace_bytes", _l), ("logits", _p)])


SIGNATURES["vidi_text_pass"] = [C.POINTER(VidiTextPass), _p]
EXTRA_SYMBOLS = ["vidi_last_error", "vidi_abi_version", "vidi_launch_count", "vidi_reset_launch_count", "vidi_text_pass_workspace_bytes"]

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once) and attach signatures.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"vidi_b200: CUDA library not found at {LIB_PATH}. Build it with `make -C vidi_b200/csrc` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    lib.vidi_last_error.restype = C.c_char_p
    lib.vidi_abi_version.restype = C.c_int
    lib.vidi_launch_count.restype = C.c_int64
    lib.vidi_reset_launch_count.restype = None
    lib.vidi_text_pass_workspace_bytes.argtypes = [C.POINTER(VidiTextPass)]
    lib.vidi_text_pass_workspace_bytes.restype = C.c_int64
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().vidi_last_error()
        raise RuntimeError(f"vidi_b200.{what} failed (rc={rc}): {msg.decode() if msg else '?'}")
