"""ctypes loader for ``libflite_b200.so`` (the C ABI declared in ``include/flite_b200.h``).

There is no fallback: if the shared library is missing or the device is not sm_100 the ops raise.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libflite_b200.so")

FLITE_OK = 0
EPI_STORE, EPI_GATED_RES, EPI_SWIGLU, EPI_QKV_ROPE = 0, 1, 2, 3
GEMM_AUTO, GEMM_1CTA_N256, GEMM_2CTA_N256, GEMM_1CTA_N128, GEMM_1CTA_N64, GEMM_GEMV = 0, 1, 2, 3, 4, 5
ATTN_AUTO, ATTN_1WG, ATTN_2WG, ATTN_2CTA_1WG, ATTN_2CTA_2WG, ATTN_2CTA_1WG_PTMEM, ATTN_2CTA_2WG_PTMEM, ATTN_QTMEM_1WG, ATTN_QTMEM_2WG, ATTN_XRES, ATTN_PERSISTENT = 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10

_P, _I, _L, _F = c_void_p, c_int, c_int64, c_float

# name -> argtypes, mirrors include/flite_b200.h one to one
SIGNATURES = {
    "flite_abi_version": [],
    "flite_last_error": [],
    "flite_check_device": [],
    "flite_set_tuning": [_I, _I],
    "flite_watchdog_status": [_P],
    "flite_cfg_euler": [_P, _I, _P, _P, _F, _F, _I, _P, _L, _P],
    "flite_apg_workspace_bytes": [],
    "flite_apg_euler": [_P, _I, _P, _P, _F, _F, _F, _P, _L, _P, _P],
    "flite_latent_unscale": [_P, _P, _F, _F, _L, _P],
    "flite_image_to_uint8": [_P, _I, _P, _I, _I, _I, _I, _P],
    "flite_groupnorm_partials_bytes": [_I, _I, _I],
    "flite_bias_residual_add_nhwc": [_P, _P, _P, _L, _I, _P],
    "flite_upsample_nearest2x_nhwc": [_P, _P, _I, _I, _I, _I, _P],
    "flite_groupnorm_silu_nhwc": [_P, _P, _P, _P, _I, _L, _I, _I, _F, _I, _P, _I, _P],
    "flite_rmsnorm_modulate": [_P, _L, _P, _L, _P, _I, _P, _P, _L, _I, _I, _I, _F, _P],
    "flite_rope_qknorm": [_P, _L, _I, _I, _P, _P, _I, _F, _P],
    "flite_patch_embed": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "flite_patch_gather": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "flite_get_tuning": [_I],
    "flite_permute_021": [_P, _P, _I, _I, _I, _P],
    "flite_timestep_embed": [_P, _I, _P, _P, _I, _I, _P],
    "flite_unpatchify": [_P, _L, _P, _I, _I, _I, _I, _I, _I, _P],
    "flite_pack_context": [_P, _L, _P, _L, _P, _I, _I, _I, _P, _P, _P, _P],
    "flite_gemm_bf16": [_P, _L, _P, _L, _P, _L, _I, _I, _I, _P, _I, _I, _P, _L, _P, _L, _I, _P, _P, _I, _F,
                        _I, _I, _I, _P],
    "flite_attention_varlen": [_P, _L, _L, _I, _P, _L, _L, _I, _P, _L, _I, _P, _L, _P, _P, _I, _I, _I, _F, _I, _P],
    "flite_attention_streamk_workspace_bytes": [],
    "flite_attention_streamk": [_P, _L, _L, _I, _P, _L, _L, _I, _P, _L, _I, _P, _L, _P, _P, _I, _I, _I, _I, _F, _P, _L, _P],
    "flite_attention_streamk_p2p": [_P, _L, _L, _I, _P, _L, _L, _I, _P, _L, _I, _P, _I, _I, _I, _L, _P, _P, _I, _I, _I, _I, _F,
                                    _P, _L, _P],
    "flite_gemm_qkv_p2p": [_P, _L, _P, _L, _I, _I, _P, _I, _P, _P, _F, _I, _I, _I, _I, _P, _I, _P],
    "flite_attention_varlen_p2p": [_P, _L, _L, _I, _P, _L, _L, _I, _P, _L, _I, _P, _I, _I, _I, _L, _P, _P, _I, _I, _I,
                                   _F, _I, _P],
    "flite_p2p_alloc": [_L, _P],
    "flite_p2p_free": [_P],
    "flite_ipc_get_handle": [_P, _P],
    "flite_ipc_open": [_P, _P],
    "flite_ipc_close": [_P],
    "flite_p2p_signal": [_P, _I, _I, ctypes.c_uint, _P],
    "flite_p2p_wait": [_P, _I, ctypes.c_uint, _P],
    "flite_poison_on_abort": [_P, _L, _P],
}

_lib = None


class FliteError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load the library (once). Raises FliteError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FliteError(
            f"{LIB_PATH} not found: build it with `python __graft_entry__.py build` "
            "(nvcc -gencode arch=compute_100a,code=sm_100a). There is no CPU / PyTorch fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = (c_char_p if name == "flite_last_error"
                      else c_int64 if name in ("flite_attention_streamk_workspace_bytes", "flite_groupnorm_partials_bytes")
                      else c_int)
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != FLITE_OK:
        msg = load().flite_last_error()
        raise FliteError(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")


def watchdog_ok() -> None:
    """Synchronise and raise if any kernel barrier wait timed out."""
    code = ctypes.c_uint(0)
    check(load().flite_watchdog_status(ctypes.byref(code)), "watchdog")
