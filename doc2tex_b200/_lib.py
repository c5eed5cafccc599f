"""ctypes binding of the C-ABI library ``libd2t_b200.so`` (include/doc2tex_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There
is no CPU fallback: if the library is missing or fails to load, importing the
engine raises immediately.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libd2t_b200.so")

# Every symbol include/doc2tex_b200.h declares (tests/test_abi.py checks the list against the header).
SYMBOLS = [
    "d2t_create", "d2t_destroy", "d2t_last_error", "d2t_version", "d2t_load_tensor",
    "d2t_finalize_weights", "d2t_encode", "d2t_encoder_geometry", "d2t_decode_greedy",
    "d2t_decode_beam", "d2t_decode_attn_greedy", "d2t_decode_attn_beam", "d2t_set_option", "d2t_set_debug", "d2t_debug_tap",
    "d2t_debug_gemm", "d2t_debug_gemm_bench", "d2t_debug_conv_time", "d2t_debug_decode_time", "d2t_debug_beam_runner_up", "d2t_launch_count", "d2t_prep_measure", "d2t_prep_render",
]

PREC = {"fp32": 0, "tf32x3": 1, "bf16x3": 2, "bf16": 3}
HEAD = {"None": 0, "TFM": 1, "Attnv2": 2, "Attn": 3}


class Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "struct_size", "in_channels", "stem_channels", "hidden", "depth", "heads", "max_tokens",
        "head", "vocab", "dec_layers", "dec_heads", "dec_ff", "max_seq_len", "attn_hidden",
        "attn_kernel_dim", "attn_kernel_size", "precision", "use_graphs")]


class PrepImage(C.Structure):
    _fields_ = [("src_off", C.c_int64), ("h0", C.c_int32), ("w0", C.c_int32), ("ds", C.c_int32), ("pad_", C.c_int32)]


class PrepPlan(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "use_crop", "crop_x", "crop_y", "crop_w", "crop_h", "inverted", "vmin", "hb", "wb", "do_resize", "rh", "rw",
        "kx_off", "kx_ksize", "ky_off", "ky_ksize", "out_h", "out_w")] + [
        ("off_b", C.c_int64), ("off_t", C.c_int64), ("off_r", C.c_int64), ("dst", C.c_void_p)]


_lib = None


def load():
    """Load the shared library (once).  Raises if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "doc2tex_b200 has no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64p, fp = C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.c_void_p
    lib.d2t_create.argtypes = [C.POINTER(Config), i32, C.POINTER(vp)]
    lib.d2t_destroy.argtypes = [vp]
    lib.d2t_last_error.argtypes = [vp]
    lib.d2t_last_error.restype = C.c_char_p
    lib.d2t_version.restype = C.c_char_p
    lib.d2t_load_tensor.argtypes = [vp, C.c_char_p, vp, i64p, i32, i32]
    lib.d2t_finalize_weights.argtypes = [vp]
    lib.d2t_encode.argtypes = [vp, fp, i32, i32, i32, fp, vp]
    lib.d2t_encoder_geometry.argtypes = [vp, i32, i32] + [C.POINTER(C.c_int)] * 5
    lib.d2t_decode_greedy.argtypes = [vp, fp, i32, i32, i32, i32, vp, fp, C.POINTER(C.c_int), vp]
    lib.d2t_decode_beam.argtypes = [vp, fp, i32, i32, i32, i32, vp, vp, fp, vp, fp, C.POINTER(C.c_int), vp]
    lib.d2t_decode_attn_beam.argtypes = [vp, fp, i32, i32, i32, i32, vp, vp, fp, vp, fp, C.POINTER(C.c_int), vp]
    lib.d2t_decode_attn_greedy.argtypes = [vp, fp, i32, i32, i32, i32, vp, fp, C.POINTER(C.c_int), vp]
    lib.d2t_set_option.argtypes = [vp, C.c_char_p, i32]
    lib.d2t_set_debug.argtypes = [vp, i32]
    lib.d2t_debug_tap.argtypes = [vp, C.c_char_p, fp, i64p, i64p, vp]
    lib.d2t_debug_gemm.argtypes = [vp, fp, fp, fp, fp, fp, i32, i32, i32, i32, i32, vp]
    lib.d2t_debug_gemm_bench.argtypes = [vp, fp, fp, fp, i32, i32, i32, i32, i32, i32, C.POINTER(C.c_float), vp]
    lib.d2t_debug_conv_time.argtypes = [vp, C.POINTER(C.c_double), i64p, C.POINTER(C.c_double)]
    lib.d2t_debug_decode_time.argtypes = [vp, i32, C.POINTER(C.c_double), i64p, C.POINTER(C.c_double)]
    lib.d2t_debug_beam_runner_up.argtypes = [vp, fp]
    lib.d2t_prep_measure.argtypes = [vp, vp, vp, i32, vp, vp]
    lib.d2t_prep_render.argtypes = [vp, vp, vp, vp, i32, vp, vp, i32, C.c_float, C.c_float, vp]
    lib.d2t_launch_count.argtypes = [vp]
    lib.d2t_launch_count.restype = C.c_int64
    for name in SYMBOLS:
        if name not in ("d2t_last_error", "d2t_version", "d2t_launch_count"):
            getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib
