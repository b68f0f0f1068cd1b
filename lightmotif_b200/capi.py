"""ctypes bindings of the C ABI (include/lightmotif_b200.h).

The shared library is built in-tree by ``lightmotif_b200/build.py`` (nvcc,
sm_100a).  Importing this module FAILS LOUDLY when the library is missing:
there is no CPU fall-back anywhere in the product path.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_PKG = Path(__file__).resolve().parent
import os

LIB_PATH = Path(os.environ.get("LM_B200_LIB", _PKG / "liblightmotif_b200.so"))  # override: experiment builds

LANES = 32
OK, EMPTY = 0, 1
ERR_INVALID_ARG, ERR_INVALID_SYMBOL, ERR_WRAP, ERR_CUDA, ERR_OOM, ERR_UNSUPPORTED, ERR_OVERFLOW = (
    -1, -2, -3, -4, -5, -6, -7)
DNA, PROTEIN = 0, 1
RULE_GENERIC, RULE_AVX2, RULE_SSE2 = 0, 1, 2
STRAND_DIRECT, STRAND_REVERSE, STRAND_BOTH = 1, 2, 3


class Hit(C.Structure):
    _fields_ = [("row", C.c_uint64), ("position", C.c_uint64), ("col", C.c_uint32), ("score", C.c_float)]


class ColWin(C.Structure):
    _fields_ = [("row", C.c_uint64), ("score", C.c_float), ("valid", C.c_uint32)]


class FilterPlan(C.Structure):
    _fields_ = [("ok", C.c_int), ("r", C.c_int), ("pair", C.c_int), ("pair_mul", C.c_uint32), ("j0", C.c_int),
                ("smax", C.c_int), ("tab_lo", C.c_uint32 * 16), ("tab_hi", C.c_uint32 * 16), ("c0", C.c_double),
                ("delta", C.c_double), ("eps", C.c_double), ("est_pass_rate", C.c_double), ("est_cost", C.c_double)]


class LightmotifError(RuntimeError):
    def __init__(self, status, detail=""):
        self.status = status
        msg = lib().lm_strerror(status).decode()
        super().__init__(f"{msg} (status {status})" + (f": {detail}" if detail else ""))


_lib = None

# every symbol declared in include/lightmotif_b200.h
SYMBOLS = [
    "lm_abi_version", "lm_strerror", "lm_ctx_last_error", "lm_ctx_create", "lm_ctx_create_on_stream",
    "lm_ctx_destroy", "lm_ctx_stream", "lm_ctx_synchronize", "lm_ctx_kernel_launches",
    "lm_ctx_last_kernel_ms", "lm_ctx_last_sm_clock_mhz", "lm_ctx_set_option", "lm_ctx_last_stats", "lm_ctx_last_upload", "lm_encode", "lm_encode_lossy", "lm_stripe", "lm_configure_wrap",
    "lm_score_f32_rows", "lm_score_u8_rows", "lm_argmax_f32", "lm_max_f32", "lm_max_f32_rule", "lm_argmax_u8", "lm_max_u8",
    "lm_threshold_f32", "lm_threshold_u8", "lm_seq_from_ascii", "lm_seq_from_symbols",
    "lm_seq_from_striped", "lm_seq_from_striped_slab", "lm_seq_from_device", "lm_seq_synthetic", "lm_seq_configure_wrap",
    "lm_seq_info", "lm_seq_device_ptr", "lm_seq_download", "lm_seq_free", "lm_seq_score_f32",
    "lm_seq_score_u8", "lm_seq_score_argmax_f32", "lm_seq_score_argmax_f32_launch",
    "lm_seq_score_argmax_f32_finish", "lm_seq_score_threshold_f32",
    "lm_seq_score_argmax_f32_batch", "lm_seq_score_threshold_f32_batch", "lm_seq_score_argmax_f32_batch_rc",
    "lm_seq_score_threshold_f32_batch_rc", "lm_pssm_reverse_complement", "lm_argmax_merge", "lm_ctx_result_block", "lm_decode_keys",
    "lm_plan_filter", "lm_plan_filter_protein",
    "lm_comm_init", "lm_comm_init_rank", "lm_comm_connect", "lm_comm_destroy", "lm_comm_last_error",
    "lm_comm_world", "lm_comm_n_local", "lm_comm_ctx", "lm_comm_rank", "lm_comm_shard_from_striped",
    "lm_comm_seq_synthetic", "lm_comm_score_argmax_f32_launch", "lm_comm_score_argmax_f32_finish",
    "lm_seq_score_argmax_f32_sharded", "lm_seq_score_threshold_f32_sharded",
]
MAX_RANKS, COMM_HANDLE_BYTES = 16, 64


def lib():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m lightmotif_b200.build` "
                "(nvcc, sm_100a). lightmotif_b200 has no CPU fall-back.")
        _lib = C.CDLL(str(LIB_PATH))
        _declare(_lib)
    return _lib


def _declare(L):
    sz, vp, i = C.c_size_t, C.c_void_p, C.c_int
    u8p, f32p = C.POINTER(C.c_uint8), C.POINTER(C.c_float)
    szp = C.POINTER(sz)
    hitp, cwp = C.POINTER(Hit), C.POINTER(ColWin)
    L.lm_abi_version.restype = i
    L.lm_strerror.argtypes = [i]
    L.lm_strerror.restype = C.c_char_p
    L.lm_ctx_last_error.argtypes = [vp]
    L.lm_ctx_last_error.restype = C.c_char_p
    L.lm_ctx_create.argtypes = [i, C.POINTER(vp)]
    L.lm_ctx_create_on_stream.argtypes = [i, vp, C.POINTER(vp)]
    L.lm_ctx_destroy.argtypes = [vp]
    L.lm_ctx_destroy.restype = None
    L.lm_ctx_stream.argtypes = [vp]
    L.lm_ctx_stream.restype = vp
    L.lm_ctx_synchronize.argtypes = [vp]
    L.lm_ctx_kernel_launches.argtypes = [vp]
    L.lm_ctx_kernel_launches.restype = C.c_uint64
    L.lm_ctx_last_kernel_ms.argtypes = [vp]
    L.lm_ctx_last_kernel_ms.restype = C.c_float
    L.lm_ctx_last_sm_clock_mhz.argtypes = [vp]
    L.lm_ctx_last_sm_clock_mhz.restype = C.c_float
    L.lm_ctx_set_option.argtypes = [vp, i, i]
    L.lm_ctx_last_stats.argtypes = [vp, C.POINTER(C.c_uint64 * 4)]
    L.lm_ctx_last_upload.argtypes = [vp, C.POINTER(C.c_uint64 * 4)]
    L.lm_encode.argtypes = [vp, i, vp, sz, vp, szp]
    L.lm_encode_lossy.argtypes = [vp, i, vp, sz, vp]
    L.lm_stripe.argtypes = [vp, i, vp, sz, vp, sz]
    L.lm_configure_wrap.argtypes = [vp, i, vp, sz, sz, sz]
    L.lm_score_f32_rows.argtypes = [vp, i, vp, sz, sz, vp, sz, sz, sz, sz, sz, sz, vp, sz, szp]
    L.lm_score_u8_rows.argtypes = [vp, i, vp, sz, sz, vp, sz, sz, sz, sz, sz, sz, vp, sz, szp]
    L.lm_argmax_f32.argtypes = [vp, vp, sz, sz, i, hitp]
    L.lm_argmax_u8.argtypes = [vp, vp, sz, sz, i, hitp]
    L.lm_max_f32.argtypes = [vp, vp, sz, sz, f32p]
    L.lm_max_f32_rule.argtypes = [vp, vp, sz, sz, i, f32p]
    L.lm_max_u8.argtypes = [vp, vp, sz, sz, u8p]
    L.lm_threshold_f32.argtypes = [vp, vp, sz, sz, C.c_float, hitp, sz, szp]
    L.lm_threshold_u8.argtypes = [vp, vp, sz, sz, C.c_uint8, hitp, sz, szp]
    L.lm_seq_from_ascii.argtypes = [vp, i, vp, sz, i, sz, C.POINTER(vp), szp]
    L.lm_seq_from_symbols.argtypes = [vp, i, vp, sz, sz, C.POINTER(vp)]
    L.lm_seq_from_striped.argtypes = [vp, i, vp, sz, sz, sz, sz, C.POINTER(vp)]
    L.lm_seq_from_striped_slab.argtypes = [vp, i, vp, sz, sz, sz, sz, C.c_uint64, C.c_uint64, C.POINTER(vp)]
    L.lm_seq_from_device.argtypes = [vp, i, vp, sz, sz, sz, C.POINTER(vp)]
    L.lm_seq_synthetic.argtypes = [vp, i, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, sz, C.POINTER(vp)]
    L.lm_seq_configure_wrap.argtypes = [vp, vp, sz]
    L.lm_seq_info.argtypes = [vp, C.POINTER(i), szp, szp, szp]
    L.lm_seq_device_ptr.argtypes = [vp]
    L.lm_seq_device_ptr.restype = vp
    L.lm_seq_download.argtypes = [vp, vp, vp, sz]
    L.lm_seq_free.argtypes = [vp, vp]
    L.lm_seq_free.restype = None
    L.lm_seq_score_f32.argtypes = [vp, vp, vp, sz, sz, sz, sz, vp, sz, i, szp]
    L.lm_seq_score_u8.argtypes = [vp, vp, vp, sz, sz, sz, sz, vp, sz, i, szp]
    L.lm_seq_score_argmax_f32.argtypes = [vp, vp, vp, sz, sz, sz, sz, i, hitp, cwp, f32p]
    L.lm_seq_score_argmax_f32_launch.argtypes = [vp, vp, vp, sz, sz, sz, sz, i]
    L.lm_seq_score_argmax_f32_finish.argtypes = [vp, i, i, hitp, cwp, f32p]
    L.lm_seq_score_threshold_f32.argtypes = [vp, vp, vp, sz, sz, sz, sz, C.c_float, hitp, sz, szp]
    L.lm_seq_score_argmax_f32_batch.argtypes = [vp, vp, vp, vp, vp, sz, sz, i, hitp]
    L.lm_seq_score_threshold_f32_batch.argtypes = [vp, vp, vp, vp, vp, sz, sz, vp, hitp, sz, vp]
    L.lm_seq_score_argmax_f32_batch_rc.argtypes = [vp, vp, vp, vp, vp, sz, sz, i, i, hitp, vp]
    L.lm_seq_score_threshold_f32_batch_rc.argtypes = [vp, vp, vp, vp, vp, sz, sz, i, vp, hitp, sz, vp]
    L.lm_pssm_reverse_complement.argtypes = [vp, sz, sz, vp]
    L.lm_argmax_merge.argtypes = [cwp, sz, vp, C.c_float, C.c_uint64, i, hitp]
    L.lm_ctx_result_block.argtypes = [vp, i]
    L.lm_ctx_result_block.restype = vp
    L.lm_decode_keys.argtypes = [vp, cwp]
    L.lm_decode_keys.restype = None
    L.lm_plan_filter.argtypes = [vp, sz, sz, C.c_double, i, C.c_double, i, i, C.POINTER(FilterPlan)]
    L.lm_plan_filter_protein.argtypes = [vp, sz, sz, C.c_double, i, C.c_double, i, C.POINTER(FilterPlan), vp]
    L.lm_comm_init.argtypes = [i, vp, sz, C.POINTER(vp)]
    L.lm_comm_init_rank.argtypes = [vp, i, i, sz, vp, C.POINTER(vp)]
    L.lm_comm_connect.argtypes = [vp, vp]
    L.lm_comm_destroy.argtypes = [vp]
    L.lm_comm_destroy.restype = None
    L.lm_comm_last_error.argtypes = [vp]
    L.lm_comm_last_error.restype = C.c_char_p
    L.lm_comm_world.argtypes = [vp]
    L.lm_comm_n_local.argtypes = [vp]
    L.lm_comm_ctx.argtypes = [vp, i]
    L.lm_comm_ctx.restype = vp
    L.lm_comm_rank.argtypes = [vp, i]
    L.lm_comm_shard_from_striped.argtypes = [vp, i, vp, sz, sz, sz, sz, C.POINTER(vp)]
    L.lm_comm_seq_synthetic.argtypes = [vp, i, C.c_uint64, C.c_uint64, sz, C.POINTER(vp)]
    L.lm_comm_score_argmax_f32_launch.argtypes = [vp, C.POINTER(vp), vp, sz, sz, i]
    L.lm_comm_score_argmax_f32_finish.argtypes = [vp, i, i, hitp]
    L.lm_seq_score_argmax_f32_sharded.argtypes = [vp, C.POINTER(vp), vp, sz, sz, i, hitp]
    L.lm_seq_score_threshold_f32_sharded.argtypes = [vp, C.POINTER(vp), vp, sz, sz, C.c_float, hitp, sz, szp, szp]
    for name in SYMBOLS:
        fn = getattr(L, name)
        if fn.restype is C.c_int and name not in ("lm_abi_version",):
            pass
