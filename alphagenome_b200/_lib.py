"""ctypes binding of libagb200.so (the C ABI in include/agb200.h).

The product path has no CPU or PyTorch fallback: if the shared library is missing, or the device is not an
sm_100 part, every op raises.  torch is used only for device memory, streams and distributed plumbing.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os

_PKG = Path(__file__).resolve().parent
# AGB200_LIB: load another build of the library (A/B timing of kernel variants on one box); default = the in-tree build
LIB_PATH = Path(os.environ["AGB200_LIB"]) if os.environ.get("AGB200_LIB") else _PKG / "libagb200.so"


class AgEpiOut(C.Structure):
    _fields_ = [
        ("ptr", C.c_void_p),
        ("scale", C.c_void_p),
        ("shift", C.c_void_p),
        ("bstride", C.c_int64),
        ("shift_bstride", C.c_int64),
        ("ld", C.c_int32),
        ("act", C.c_int32),
        ("dtype", C.c_int32),
        ("reserved", C.c_int32),
    ]


class AgGemmArgs(C.Structure):
    _fields_ = [
        ("A", C.c_void_p),
        ("a_bstride", C.c_int64),
        ("a_ld", C.c_int32),
        ("a_rows", C.c_int32),
        ("a_row0", C.c_int32),
        ("w_batched", C.c_int32),
        ("W", C.c_void_p),
        ("w_zstride", C.c_int64),
        ("w_ld", C.c_int32),
        ("batch", C.c_int32),
        ("M", C.c_int32),
        ("N", C.c_int32),
        ("K", C.c_int32),
        ("taps", C.c_int32),
        ("w_rows", C.c_int32),
        ("cta_group", C.c_int32),
        ("pre_scale", C.c_void_p),
        ("pre_shift", C.c_void_p),
        ("relu", C.c_int32),
        ("res_ch", C.c_int32),
        ("res", C.c_void_p),
        ("res_bstride", C.c_int64),
        ("res_ld", C.c_int32),
        ("res_shift", C.c_int32),
        ("post_scale", C.c_void_p),
        ("out", C.c_void_p),
        ("out_bstride", C.c_int64),
        ("out_ld", C.c_int32),
        ("pool_ld", C.c_int32),
        ("pool", C.c_void_p),
        ("pool_bstride", C.c_int64),
        ("actA", AgEpiOut),
        ("actB", AgEpiOut),
        ("actP", AgEpiOut),
        ("A2", C.c_void_p),
        ("W2", C.c_void_p),
        ("a2_bstride", C.c_int64),
        ("a2_ld", C.c_int32),
        ("a2_rows", C.c_int32),
        ("a2_row0", C.c_int32),
        ("w2_ld", C.c_int32),
        ("K2", C.c_int32),
        ("reserved2", C.c_int32),
    ]


class AgAttnArgs(C.Structure):
    _fields_ = [
        ("Q", C.c_void_p),
        ("K", C.c_void_p),
        ("Vt", C.c_void_p),
        ("q_bstride", C.c_int64),
        ("q_hstride", C.c_int64),
        ("k_bstride", C.c_int64),
        ("k_hstride", C.c_int64),
        ("v_bstride", C.c_int64),
        ("v_hstride", C.c_int64),
        ("q_ld", C.c_int32),
        ("k_ld", C.c_int32),
        ("v_ld", C.c_int32),
        ("batch", C.c_int32),
        ("heads", C.c_int32),
        ("kv_heads", C.c_int32),
        ("n_q", C.c_int32),
        ("n_k", C.c_int32),
        ("d", C.c_int32),
        ("dv", C.c_int32),
        ("mode", C.c_int32),
        ("bias_p", C.c_int32),
        ("scale", C.c_float),
        ("softcap", C.c_float),
        ("bias", C.c_void_p),
        ("out", C.c_void_p),
        ("out_bstride", C.c_int64),
        ("out_hstride", C.c_int64),
        ("out_ld", C.c_int32),
        ("out_dtype", C.c_int32),
        ("out_bias", C.c_void_p),
        ("res", C.c_void_p),
        ("bias_rows", C.c_int32),
        ("reserved1", C.c_int32),
        ("norm_w", C.c_void_p),
        ("norm_b", C.c_void_p),
        ("norm_out", C.c_void_p),
    ]


class AgAffineArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p),
        ("scale", C.c_void_p),
        ("shift", C.c_void_p),
        ("res", C.c_void_p),
        ("out_f32", C.c_void_p),
        ("out_bf16", C.c_void_p),
        ("x_bstride", C.c_int64),
        ("shift_bstride", C.c_int64),
        ("res_bstride", C.c_int64),
        ("out_f32_bstride", C.c_int64),
        ("out_bf16_bstride", C.c_int64),
        ("x_ld", C.c_int32),
        ("res_ld", C.c_int32),
        ("out_f32_ld", C.c_int32),
        ("out_bf16_ld", C.c_int32),
        ("B", C.c_int32),
        ("rows", C.c_int32),
        ("C", C.c_int32),
        ("act", C.c_int32),
    ]


ACT_NONE, ACT_GELU1702, ACT_GELU_ERF, ACT_RELU, ACT_SOFTPLUS_MUL, ACT_SIGMOID = 0, 1, 2, 3, 4, 5

# every symbol include/agb200.h declares; tests check the built library exports all of them
EXPORTED = [
    "ag_version",
    "ag_last_error",
    "ag_launch_count",
    "ag_check_device",
    "ag_watchdog_status",
    "ag_gemm_set_cta_group",
    "ag_gemm_fwd",
    "ag_attention_fwd",
    "ag_onehot_im2col_fwd",
    "ag_qkv_post_fwd",
    "ag_transpose_bf16",
    "ag_pair_bias_fwd",
    "ag_pool_rmsnorm_fwd",
    "ag_pair_combine_fwd",
    "ag_rmsnorm_fwd",
    "ag_add_embed_norm_fwd",
    "ag_pair_out_embed_fwd",
    "ag_rmsnorm_act_fwd",
    "ag_splice_rope_fwd",
    "ag_bn_stats_fwd",
    "ag_bn_update_fwd",
    "ag_affine_act_fwd",
    "ag_qkv_post_gather_fwd",
    "ag_peer_push_fwd",
]

_lib = None


class AgbError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load libagb200.so (building is `python -m alphagenome_b200.csrc.build` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise AgbError(
            f"{LIB_PATH} is missing: build it with `python -m alphagenome_b200.csrc.build` "
            "(there is no CPU/PyTorch fallback for the trunk kernels)"
        )
    lib = C.CDLL(str(LIB_PATH))
    lib.ag_version.restype = C.c_int
    lib.ag_last_error.restype = C.c_char_p
    lib.ag_launch_count.restype = C.c_uint64
    lib.ag_check_device.argtypes = [C.c_int]
    lib.ag_check_device.restype = C.c_int
    lib.ag_watchdog_status.argtypes = [C.c_int, C.POINTER(C.c_uint32 * 4)]
    lib.ag_watchdog_status.restype = C.c_int
    lib.ag_gemm_fwd.argtypes = [C.POINTER(AgGemmArgs), C.c_int, C.c_void_p]
    lib.ag_gemm_fwd.restype = C.c_int
    P, I32, I64 = C.c_void_p, C.c_int32, C.c_int64
    lib.ag_attention_fwd.argtypes = [C.POINTER(AgAttnArgs), C.c_int, P]
    lib.ag_onehot_im2col_fwd.argtypes = [P, P, I32, I32, I32, I32, I32, C.c_int, P]
    lib.ag_qkv_post_fwd.argtypes = [P] * 12 + [I32] * 5 + [C.c_int, P]
    lib.ag_transpose_bf16.argtypes = [P, P, I32, I32, I32, I64, I32, I64, I32, C.c_int, P]
    lib.ag_pair_bias_fwd.argtypes = [P] * 5 + [I32] * 6 + [I64, C.c_int, P]
    lib.ag_pool_rmsnorm_fwd.argtypes = [P] * 5 + [I32] * 4 + [C.c_int, P]
    lib.ag_pair_combine_fwd.argtypes = [P] * 11 + [I32] * 8 + [C.c_int, P]
    lib.ag_rmsnorm_fwd.argtypes = [P] * 4 + [I64, I32, C.c_int, P]
    lib.ag_add_embed_norm_fwd.argtypes = [P] * 6 + [I32] * 3 + [C.c_int, P]
    lib.ag_pair_out_embed_fwd.argtypes = [P] * 6 + [I32] * 4 + [C.c_int, P]
    lib.ag_rmsnorm_act_fwd.argtypes = [P] * 5 + [I64, I64, I32, I32, C.c_int, P]
    lib.ag_splice_rope_fwd.argtypes = [P] * 6 + [I32] * 6 + [C.c_int, P]
    lib.ag_qkv_post_gather_fwd.argtypes = [P, P, C.POINTER(C.c_void_p), I32, I64, I64] + [P] * 8 + [I32] * 5 + [C.c_int, P]
    lib.ag_peer_push_fwd.argtypes = [P, C.POINTER(C.c_void_p), I32, I64, I64, C.c_int, P]
    lib.ag_bn_stats_fwd.argtypes = [P, P, I32, I32, I64, I32, I32, C.c_int, P]
    lib.ag_bn_update_fwd.argtypes = [P, C.c_double, P, P, C.c_float, C.c_float, P, P, P, P, I32, C.c_int, P]
    lib.ag_affine_act_fwd.argtypes = [C.POINTER(AgAffineArgs), C.c_int, P]
    lib.ag_gemm_set_cta_group.argtypes = [C.c_int]
    lib.ag_gemm_set_cta_group.restype = C.c_int
    for name in EXPORTED[6:]:
        getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib


def last_error() -> str:
    return load().ag_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != 0:
        raise AgbError(last_error())


def launch_count() -> int:
    return int(load().ag_launch_count())


def watchdog_status(device: int = 0):
    out = (C.c_uint32 * 4)()
    load().ag_watchdog_status(device, C.byref(out))
    return tuple(int(x) for x in out)
