"""ctypes binding of `libnumpitron_b200.so` (the C ABI in include/numpitron_b200.h).

There is NO fallback: if the shared object is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "lib" / "libnumpitron_b200.so"

c_void_p, c_int, c_int32, c_int64, c_float, c_size_t = C.c_void_p, C.c_int, C.c_int32, C.c_int64, C.c_float, C.c_size_t


class NptError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    """Mirror of `npt_gemm_args`."""

    _fields_ = [
        ("mode", c_int32), ("trans_a", c_int32), ("trans_b", c_int32),
        ("M", c_int32), ("N", c_int32), ("K", c_int32),
        ("batch", c_int32), ("batch_inner", c_int32),
        ("A", c_void_p), ("lda", c_int64), ("stride_a_outer", c_int64), ("stride_a_inner", c_int64),
        ("B", c_void_p), ("ldb", c_int64), ("stride_b_outer", c_int64), ("stride_b_inner", c_int64),
        ("C", c_void_p), ("ldc", c_int64), ("stride_c_outer", c_int64), ("stride_c_inner", c_int64),
        ("C_bf16", c_void_p), ("ldcb", c_int64), ("stride_cb_outer", c_int64), ("stride_cb_inner", c_int64),
        ("alpha", c_float),
        ("bias", c_void_p),
        ("relu", c_int32),
        ("mask", c_void_p),
        ("mask_is_bf16", c_int32),
        ("ldmask", c_int64), ("stride_mask_outer", c_int64), ("stride_mask_inner", c_int64),
        ("workspace", c_void_p), ("workspace_bytes", c_size_t),
        ("relu_bits_out", c_void_p), ("mask_bits", c_void_p), ("ld_bits", c_int64),
        ("C_bf16_peer", c_void_p),
        ("peer_rows_lo", c_int32), ("peer_rows_hi", c_int32),
        ("skip_splitk_reduce", c_int32),
    ]


class AdamTensor(C.Structure):
    """Mirror of `npt_adam_tensor`."""

    _fields_ = [
        ("p", c_void_p), ("g", c_void_p), ("m", c_void_p), ("v", c_void_p),
        ("p_bf16", c_void_p), ("n", c_int64), ("shadow_cols", c_int64), ("shadow_ld", c_int64),
        ("g_splits", c_int64), ("g_split_stride", c_int64),
    ]


_P = c_void_p
_SIGNATURES = {
    "npt_last_error": (C.c_char_p, []),
    "npt_version": (c_int, []),
    "npt_launch_count": (C.c_ulonglong, []),
    "npt_set_sm_margin": (c_int, [c_int32]),
    "npt_device_info": (c_int, [C.POINTER(c_int), C.POINTER(c_int), C.POINTER(c_int)]),
    "npt_pinned_alloc": (c_int, [C.POINTER(c_void_p), c_size_t]),
    "npt_pinned_free": (c_int, [_P]),
    "npt_memcpy_async": (c_int, [_P, _P, c_size_t, c_int32, _P]),
    "npt_peer_barrier": (c_int, [_P, _P, _P, _P]),
    "npt_memcpy2d_async": (c_int, [_P, c_size_t, _P, c_size_t, c_size_t, c_size_t, _P]),
    "npt_gemm_workspace_bytes": (c_size_t, [C.POINTER(GemmArgs)]),
    "npt_gemm": (c_int, [C.POINTER(GemmArgs), _P]),
    "npt_cast_f32_bf16": (c_int, [_P, c_int64, _P, c_int64, c_int64, c_int64, _P]),
    "npt_layernorm_fwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int32, c_float, _P]),
    "npt_layernorm_bwd_workspace_bytes": (c_size_t, [c_int64, c_int32]),
    "npt_layernorm_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int32, c_float, _P, c_size_t, _P]),
    "npt_layernorm_bwd_colsum_supported": (c_int, [c_int32]),
    "npt_gemm_splits": (c_int32, [_P]),
    "npt_window_gather": (c_int, [_P, c_int64, _P, _P, _P, c_int32, c_int32, _P]),
    "npt_layernorm_bf16_input_supported": (c_int, [c_int32]),
    "npt_layernorm_fwd_v2": (c_int, [_P, c_int32, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int32, c_float, _P]),
    "npt_layernorm_bwd_v2": (c_int, [_P, _P, c_int32, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int32, c_float, _P, c_size_t, _P]),
    "npt_layernorm_fwd_v3": (c_int, [_P, c_int32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int32, c_float, _P]),
    "npt_layernorm_bwd_v4": (c_int, [_P, _P, c_int32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int32, c_float, _P, c_size_t,
                                     _P, c_int32, _P, _P]),
    "npt_layernorm_bwd_partials": (c_int32, [c_int64, c_int32, c_int32]),
    "npt_layernorm_bwd_reduce": (c_int, [_P, c_int32, c_int32, c_int32, _P, _P, _P, _P]),
    "npt_layernorm_bwd_v3": (c_int, [_P, _P, c_int32, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int32, c_float, _P, c_size_t,
                                     c_int32, _P, _P]),
    "npt_relu_fwd": (c_int, [_P, _P, _P, c_int64, _P]),
    "npt_relu_bwd": (c_int, [_P, _P, _P, _P, c_int64, _P]),
    "npt_colsum_workspace_bytes": (c_size_t, [c_int64, c_int32]),
    "npt_colsum": (c_int, [_P, c_int32, c_int64, _P, c_int64, c_int32, _P, c_size_t, _P]),
    "npt_softmax_causal_fwd": (c_int, [_P, _P, _P, c_int64, c_int32, c_float, _P]),
    "npt_softmax_causal_bwd": (c_int, [_P, c_int32, _P, _P, _P, c_int64, c_int32, c_float, _P]),
    "npt_attention_supported": (c_int, [c_int32, c_int32]),
    "npt_attention_fwd": (c_int, [_P, _P, _P, c_int32, c_int32, c_int32, c_int32, c_float, _P]),
    "npt_attention_bwd_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32]),
    "npt_attention_bwd": (c_int, [_P, _P, _P, _P, _P, c_int32, c_int32, c_int32, c_int32, c_float, _P, c_size_t, _P]),
    "npt_embedding_gather": (c_int, [_P, _P, c_int64, _P, _P, _P, c_int64, c_int32, c_int32, c_int32, c_int32, c_int32, _P]),
    "npt_add_pos": (c_int, [_P, _P, _P, _P, c_int64, c_int32, c_int32, _P]),
    "npt_embedding_scatter_add_workspace_bytes": (c_size_t, [c_int64, c_int32]),
    "npt_embedding_scatter_add": (c_int, [_P, _P, _P, c_int64, c_int64, c_int32, c_int32, c_int32, c_int32, c_int32, _P, c_size_t, _P]),
    "npt_softmax_ce_stage1": (c_int, [_P, c_int64, _P, c_int64, c_int32, _P]),
    "npt_softmax_ce_stage2": (c_int, [_P, c_int64, _P, _P, _P, _P, c_int64, c_int32, c_int32, c_int32, _P]),
    "npt_softmax_ce_stage3": (c_int, [_P, c_int64, _P, _P, _P, _P, _P, _P, c_int64, _P, c_int64, c_int64, c_int32, c_int32, c_int32, _P]),
    "npt_softmax_ce_fused": (c_int, [_P, c_int64, _P, _P, _P, c_int64, _P, c_int64, c_int64, c_int32, c_int32, c_int32, _P]),
    "npt_mean": (c_int, [_P, _P, c_int64, _P]),
    "npt_adam_step": (c_int, [_P, C.POINTER(AdamTensor), c_int32, c_float, c_float, c_float, c_float, c_float, c_float, c_float, c_float, _P]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared object (once).  Raises NptError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise NptError(
            f"{LIB_PATH} is missing: numpitron_b200 has no CPU/PyTorch fallback. "
            "Build it with `python -m numpitron_b200.build` (nvcc, sm_100a).")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def declared_symbols():
    return list(_SIGNATURES)


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().npt_last_error().decode(errors="replace")
        raise NptError(f"{what or 'numpitron_b200 call'} failed (rc={rc}): {msg}")
