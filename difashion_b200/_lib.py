"""ctypes binding of ``libdfb200.so`` (the C ABI declared in ``include/dfb200.h``).

There is deliberately no fallback: if the library cannot be loaded every op raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
# DFB200_LIB=<path>: load another build of the library (A/B runs of kernel variants on one GPU box)
LIB_PATH = os.environ.get("DFB200_LIB") or os.path.join(_HERE, "libdfb200.so")

DTYPE_BF16 = 0
DTYPE_F32 = 1


class DfbError(RuntimeError):
    pass


class GemmParams(C.Structure):
    """Mirror of ``dfb_gemm_params`` (include/dfb200.h)."""
    _fields_ = [
        ("a", C.c_void_p * 2),
        ("a_ld", C.c_int32 * 2),
        ("a_c", C.c_int32 * 2),
        ("ntaps", C.c_int32 * 2),
        ("tap_dh", (C.c_int32 * 9) * 2),
        ("tap_dw", (C.c_int32 * 9) * 2),
        ("tap_coff", (C.c_int32 * 9) * 2),
        ("nseg", C.c_int32),
        ("conv", C.c_int32),
        ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("M", C.c_int32), ("N", C.c_int32),
        ("w", C.c_void_p),
        ("w_ld", C.c_int32),
        ("bias", C.c_void_p),
        ("rowbias", C.c_void_p),
        ("rowbias_ld", C.c_int32),
        ("rows_per_batch", C.c_int32),
        ("residual", C.c_void_p),
        ("res_ld", C.c_int32),
        ("res_dtype", C.c_int32),
        ("out", C.c_void_p),
        ("out_ld", C.c_int32),
        ("out_dtype", C.c_int32),
        ("geglu", C.c_int32),
        ("act", C.c_int32),
        ("block_n", C.c_int32),
        ("gn_partial", C.c_void_p),
        ("cta_group", C.c_int32),
        ("up2x", C.c_int32),
    ]


class AttnParams(C.Structure):
    """Mirror of ``dfb_attn_params`` (include/dfb200.h)."""
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("out", C.c_void_p),
        ("q_ld", C.c_int32), ("k_ld", C.c_int32), ("v_ld", C.c_int32), ("out_ld", C.c_int32),
        ("q_col0", C.c_int32), ("k_col0", C.c_int32), ("v_col0", C.c_int32), ("out_col0", C.c_int32),
        ("B", C.c_int32), ("heads", C.c_int32), ("Sq", C.c_int32), ("Skv", C.c_int32), ("dp", C.c_int32),
        ("scale", C.c_float),
        ("block_kv", C.c_int32),
        ("dbg_v_lbo", C.c_int32), ("dbg_v_sbo", C.c_int32),
        ("dbg_flags", C.c_int32),
        ("dbg_timeline", C.c_void_p),
        ("causal", C.c_int32),
        ("ones_col", C.c_int32),
        ("workspace", C.c_void_p),
    ]


_lib: Optional[C.CDLL] = None

# name -> (restype, argtypes).  Every symbol include/dfb200.h declares is listed here; the
# CPU test-suite checks that the built library exports all of them.
_PROTOTYPES = {
    "dfb_strerror": (C.c_char_p, [C.c_int]),
    "dfb_last_error": (C.c_char_p, []),
    "dfb_abi_version": (C.c_int, []),
    "dfb_num_sms": (C.c_int, []),
    "dfb_sizeof_gemm_params": (C.c_size_t, []),
    "dfb_attention_ws_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "dfb_sizeof_attn_params": (C.c_size_t, []),
    "dfb_gemm": (C.c_int, [C.POINTER(GemmParams), C.c_void_p]),
    "dfb_gemm_f32": (C.c_int, [C.POINTER(GemmParams), C.c_void_p]),
    "dfb_geglu_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "dfb_attention": (C.c_int, [C.POINTER(AttnParams), C.c_void_p]),
    "dfb_attention_f32": (C.c_int, [C.POINTER(AttnParams), C.c_void_p]),
    "dfb_groupnorm_ws_floats": (C.c_size_t, [C.c_int, C.c_int]),
    "dfb_groupnorm": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "dfb_groupnorm_fused": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                      C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "dfb_layernorm": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int, C.c_int,
                                C.c_int, C.c_int, C.c_void_p]),
    "dfb_softmax_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_void_p]),
    "dfb_cfg_step": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_void_p, C.c_float,
                               C.POINTER(C.c_float), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
                               C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "dfb_mutual_gather_sum": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_int, C.c_void_p]),
    "dfb_mutual_blend": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int,
                                   C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int, C.c_int, C.c_void_p, C.c_int,
                                   C.c_void_p]),
    "dfb_nchw_to_nhwc": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "dfb_nhwc_to_nchw": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "dfb_pad_cast_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_void_p]),
    "dfb_cast_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_void_p]),
    "dfb_upsample2x": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "dfb_space_to_depth": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "dfb_timestep_embedding": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                         C.c_void_p]),
    "dfb_embed_tokens": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_void_p]),
    "dfb_image_to_uint8": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p]),
}


def load() -> C.CDLL:
    """Load the CUDA library, failing loudly when it is missing (no CPU/eager fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DfbError(
            f"{LIB_PATH} not found: the CUDA extension is mandatory. Build it with "
            "`python -m difashion_b200.build` (or `python -c 'import __graft_entry__ as g; g.build()'`).")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        lib = load()
        raise DfbError(f"{what}: {lib.dfb_strerror(rc).decode()} — {lib.dfb_last_error().decode()}")


def exported_symbols():
    return list(_PROTOTYPES.keys())
