"""ctypes binding of the C-ABI library (include/r3d_b200.h).

There is no CPU fallback: if the shared object is missing or a tensor is not on
a CUDA device the call raises.  The library is looked up in-tree
(r3d_b200/csrc/libr3d_b200.so) so the GPU box loads exactly what was built here.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libr3d_b200.so")

_lib = None

# name -> (restype, argtypes); mirrors include/r3d_b200.h one to one
SIGNATURES = {
    "r3d_last_error": (c_char_p, []),
    "r3d_abi_version": (c_int, []),
    "r3d_launch_count": (c_int64, [c_int]),
    "r3d_set_option": (c_int, [c_char_p, ctypes.c_double]),
    "r3d_debug_panel_round": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p,
                                      c_void_p]),
    "r3d_debug_vchain": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p]),
    "r3d_debug_round_plan": (c_int, [c_int, c_void_p, c_int]),
    "r3d_debug_chain_plan": (c_int, [c_int, c_int, c_int, c_void_p]),
    "r3d_profile_enable": (c_int, [c_int]),
    "r3d_profile_num_stages": (c_int, []),
    "r3d_profile_stage_name": (c_char_p, [c_int]),
    "r3d_profile_read": (c_int, [c_void_p, c_void_p, c_void_p, c_int]),
    "r3d_score_workspace_floats": (c_size_t, [c_int64, c_int64]),
    "r3d_channel_score_partial": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p]),
    "r3d_score_finalize": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "r3d_bottomk": (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p]),
    "r3d_score_finalize_packed": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_void_p]),
    "r3d_score_pack": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_void_p]),
    "r3d_bottomk_scaled": (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "r3d_bn_workspace_floats": (c_size_t, [c_int64, c_int64]),
    "r3d_bn_stats": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "r3d_exchange_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int,
                                 c_void_p, c_int64, c_int64, c_int, c_void_p]),
    "r3d_exchange_bwd_workspace_floats": (c_size_t, [c_int64, c_int64]),
    "r3d_exchange_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                                 c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p]),
    "r3d_exchange_bwd_finalize": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "r3d_bn_bwd_apply": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_int64, c_int64, c_int, c_void_p]),
    "r3d_erank_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64, c_int]),
    "r3d_erank_fwd": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int, c_float, c_int, c_void_p, c_void_p,
                              c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "r3d_erank_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int,
                              c_float, c_void_p, c_void_p, c_int, c_void_p]),
    "r3d_gram": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "r3d_jacobi_eigh": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                c_void_p]),
    "r3d_jacobi_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "r3d_token_informativeness": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_float, c_void_p, c_void_p]),
    "r3d_gemm_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64, c_int]),
    "r3d_gemm": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_int, c_void_p, c_void_p,
                         c_void_p]),
    "r3d_colsum_finalize": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p]),
    "r3d_colsum_workspace_floats": (c_size_t, [c_int64, c_int64]),
    "r3d_colsum": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "r3d_relu_bwd": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "r3d_exchange_one_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int,
                                     c_void_p]),
    "r3d_exchange_one_bwd": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64,
                                     c_int, c_void_p]),
    "r3d_mtoken_attn_fwd": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int64, c_int, c_int, c_void_p]),
    "r3d_mtoken_attn_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int64, c_int, c_int, c_void_p]),
    "r3d_token_mean": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int64, c_int, c_int, c_void_p]),
    "r3d_token_scores": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_float, c_void_p, c_void_p]),
    "r3d_token_mask": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p]),
    "r3d_token_exchange_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p]),
    "r3d_token_exchange_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p]),
    "r3d_panel_tiles": (c_int, [c_void_p, c_int]),
    "r3d_stream_sets_created": (c_int, []),
    "r3d_ln_bwd_workspace_floats": (c_size_t, [c_int64, c_int64]),
    "r3d_ln_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_float, c_int, c_void_p, c_void_p,
                           c_void_p, c_void_p]),
    "r3d_ln_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_void_p,
                           c_void_p, c_void_p, c_void_p]),
    "r3d_ln_fwd2": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_float, c_int, c_void_p, c_void_p,
                            c_void_p, c_void_p]),
    "r3d_ln_bwd2": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_void_p,
                            c_void_p, c_void_p, c_void_p, c_void_p]),
    "r3d_ln_relu_parts": (c_int64, [c_int64]),
    "r3d_ln_relu_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_float, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p]),
    "r3d_ln_relu_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int,
                                c_void_p, c_void_p, c_void_p, c_void_p]),
    "r3d_swap_add": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p]),
    "r3d_fuser_step_host": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_int64, c_float, c_float,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "r3d_token_fusion_host": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_int64, c_void_p,
                                      c_void_p, c_void_p, c_void_p]),
}


class Epilogue(ctypes.Structure):
    """struct r3d_epilogue (include/r3d_b200.h)."""
    _fields_ = [("bias", c_void_p), ("residual", c_void_p), ("aux_out", c_void_p), ("aux_in", c_void_p),
                ("colsum_partial", c_void_p), ("act", c_int), ("colsum_abs", c_int)]


class R3DError(RuntimeError):
    pass


def lib():
    """Load (once) and return the ctypes handle.  Raises if the library was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise R3DError(
                f"{LIB_PATH} not found: build it with `python -m r3d_b200.csrc.build` "
                "(r3d_b200 has no CPU or eager fallback)")
        h = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(h, name)      # AttributeError here = header/library drift
            fn.restype = res
            fn.argtypes = args
        _lib = h
    return _lib


def check(rc: int):
    if rc != 0:
        msg = lib().r3d_last_error()
        raise R3DError(f"r3d_b200 error {rc}: {msg.decode() if msg else '?'}")


def launch_count(reset: bool = False) -> int:
    return int(lib().r3d_launch_count(1 if reset else 0))


def panel_tiles(reset: bool = True):
    """(G tiles, V tiles, group-local tiles) the Jacobi panel kernel processed on the current device since the
    last reset."""
    out = (ctypes.c_uint64 * 3)()
    check(lib().r3d_panel_tiles(out, 1 if reset else 0))
    return int(out[0]), int(out[1]), int(out[2])


def profile_enable(on: bool = True) -> bool:
    return bool(lib().r3d_profile_enable(1 if on else 0))


def profile_read(reset: bool = True) -> dict:
    """{stage: {"ms": total, "calls": n, "launches": k}} for stages that ran."""
    L = lib()
    n = L.r3d_profile_num_stages()
    ms = (ctypes.c_double * n)()
    calls = (ctypes.c_int64 * n)()
    kl = (ctypes.c_int64 * n)()
    check(L.r3d_profile_read(ms, calls, kl, 1 if reset else 0))
    out = {}
    for i in range(n):
        if calls[i]:
            out[L.r3d_profile_stage_name(i).decode()] = {"ms": ms[i], "calls": int(calls[i]), "launches": int(kl[i])}
    return out


def set_option(key: str, value: float):
    check(lib().r3d_set_option(key.encode(), float(value)))
