"""ctypes binding of libqasr.so (C ABI declared in include/qasr.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C qwen3_asr_mlx_b200/csrc``.
There is deliberately no fallback: if the shared object is missing or no sm_100 GPU is
present, the calls raise.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint16, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QASR_LIB_PATH") or os.path.join(_HERE, "lib", "libqasr.so")  # override: A/B builds

QASR_F32 = 0
QASR_BF16 = 1

QASR_OK = 0
QASR_ERR_INVALID = -1
QASR_ERR_CUDA = -2
QASR_ERR_UNSUPPORTED = -3
QASR_ERR_STATE = -4
QASR_ERR_NOMEM = -5


class QasrConfig(Structure):
    _fields_ = [
        ("d_model", c_int32),
        ("encoder_layers", c_int32),
        ("encoder_attention_heads", c_int32),
        ("encoder_ffn_dim", c_int32),
        ("num_mel_bins", c_int32),
        ("max_source_positions", c_int32),
        ("output_dim", c_int32),
        ("n_window", c_int32),
        ("n_window_infer", c_int32),
        ("downsample_hidden_size", c_int32),
    ]


class QasrStats(Structure):
    _fields_ = [("kernel_launches", c_uint64), ("workspace_bytes", c_uint64), ("weight_bytes", c_uint64)]


class QasrDecoderConfig(Structure):
    _fields_ = [
        ("hidden_size", c_int32), ("num_hidden_layers", c_int32), ("num_attention_heads", c_int32), ("num_key_value_heads", c_int32),
        ("head_dim", c_int32), ("intermediate_size", c_int32), ("vocab_size", c_int32), ("rms_norm_eps", c_float), ("rope_theta", c_float),
    ]


QASR_DEVICE_PTR = 0x100
PROF_CATEGORIES = 13


class QasrProfile(Structure):
    _fields_ = [("ms", ctypes.c_double * PROF_CATEGORIES), ("flops", ctypes.c_double * PROF_CATEGORIES),
                ("bytes", ctypes.c_double * PROF_CATEGORIES), ("launches", c_uint64 * PROF_CATEGORIES)]


class QasrError(RuntimeError):
    """A libqasr call failed with a non-argument error (CUDA, state, memory, unsupported)."""


# name -> (restype, argtypes); must list every symbol include/qasr.h declares.
_SIGNATURES = {
    "qasr_default_config": (None, [POINTER(QasrConfig)]),
    "qasr_create": (c_int, [c_int, POINTER(QasrConfig), POINTER(c_void_p)]),
    "qasr_destroy": (None, [c_void_p]),
    "qasr_last_error": (c_char_p, [c_void_p]),
    "qasr_set_weight": (c_int, [c_void_p, c_char_p, c_void_p, c_int, c_int, POINTER(c_int64)]),
    "qasr_finalize_weights": (c_int, [c_void_p]),
    "qasr_count_frames": (c_int, [c_int64, POINTER(c_int64)]),
    "qasr_count_tokens": (c_int, [c_void_p, c_int64, POINTER(c_int64)]),
    "qasr_reserve": (c_int, [c_void_p, c_int64, c_int32]),
    "qasr_mel": (c_int, [c_void_p, c_void_p, POINTER(c_int64), c_int32, c_void_p, c_void_p]),
    "qasr_encode": (c_int, [c_void_p, c_void_p, POINTER(c_int64), c_int32, c_void_p, c_int, POINTER(c_int64), c_void_p]),
    "qasr_encode_audio": (c_int, [c_void_p, c_void_p, POINTER(c_int64), c_int32, c_void_p, c_int, POINTER(c_int64), c_void_p]),
    "qasr_mel_host": (c_int, [c_void_p, c_void_p, POINTER(c_int64), c_int32, c_void_p]),
    "qasr_encode_host": (c_int, [c_void_p, c_void_p, POINTER(c_int64), c_int32, c_void_p, c_int, POINTER(c_int64)]),
    "qasr_encode_audio_hidden": (c_int, [c_void_p, c_void_p, POINTER(c_int64), c_int32, POINTER(c_int64), c_void_p]),
    "qasr_project_rows": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int, c_void_p]),
    "qasr_encode_audio_host": (c_int, [c_void_p, c_void_p, POINTER(c_int64), c_int32, c_void_p, c_int, POINTER(c_int64)]),
    "qasr_encode_audio_host_async": (c_int, [c_void_p, c_int32, c_void_p, POINTER(c_int64), c_int32, c_void_p, c_int, POINTER(c_int64)]),
    "qasr_host_wait": (c_int, [c_void_p, c_int32]),
    "qasr_prepare_inputs": (c_int, [c_void_p, POINTER(c_int32), c_int64, c_void_p, c_int, c_int64, c_int32, c_void_p, c_int, c_int64, c_int32, c_void_p, c_void_p]),
    "qasr_find_split_points": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int32, POINTER(c_int64), c_int32, POINTER(c_int32), c_void_p, c_void_p]),
    "qasr_pack_audio": (c_int, [c_void_p, POINTER(c_void_p), POINTER(c_int64), c_int32, c_void_p, c_void_p]),
    "qasr_scatter_rows_to_peers": (c_int, [c_void_p, c_int64, c_int32, c_void_p, POINTER(c_void_p), c_int32, c_void_p]),
    "qasr_mel_filterbank": (c_int, [POINTER(c_float)]),
    "qasr_hann_window": (c_int, [POINTER(c_float)]),
    "qasr_positional_embedding": (c_int, [c_void_p, c_int32, POINTER(c_float)]),
    "qasr_get_stats": (c_int, [c_void_p, POINTER(QasrStats)]),
    "qasr_set_profile": (c_int, [c_void_p, c_int]),
    "qasr_get_profile": (c_int, [c_void_p, POINTER(QasrProfile)]),
    "qasr_profile_name": (c_char_p, [c_int]),
    "qasr_set_debug": (c_int, [c_void_p, c_int]),
    "qasr_debug_read": (c_int, [c_void_p, c_char_p, POINTER(c_float), c_size_t]),
    # include/qasr_decoder.h
    "qasr_decoder_default_config": (None, [POINTER(QasrDecoderConfig)]),
    "qasr_decoder_create": (c_int, [c_int, POINTER(QasrDecoderConfig), POINTER(c_void_p)]),
    "qasr_decoder_destroy": (None, [c_void_p]),
    "qasr_decoder_last_error": (c_char_p, [c_void_p]),
    "qasr_decoder_set_weight": (c_int, [c_void_p, c_char_p, c_void_p, c_int, c_int, POINTER(c_int64)]),
    "qasr_decoder_finalize": (c_int, [c_void_p]),
    "qasr_decoder_embed_table": (c_int, [c_void_p, POINTER(c_void_p), POINTER(c_int)]),
    "qasr_decoder_prefill": (c_int, [c_void_p, c_void_p, c_int, POINTER(c_int64), c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "qasr_decoder_get_stats": (c_int, [c_void_p, POINTER(QasrStats)]),
    "qasr_bench_gemm": (c_int, [c_int, c_int32, c_int32, c_int32, c_int32, c_int32, POINTER(c_float)]),
    "qasr_test_gelu": (c_int, [c_int, POINTER(c_float), c_int32, POINTER(c_float)]),
    "qasr_test_gemm": (c_int, [c_int, POINTER(c_uint16), POINTER(c_uint16), POINTER(c_float), c_int32, c_int32, c_int32, c_int32, POINTER(c_float)]),
}

_lib = None


def exported_symbols() -> list[str]:
    return sorted(_SIGNATURES)


def load() -> ctypes.CDLL:
    """Load libqasr.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise QasrError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C qwen3_asr_mlx_b200/csrc` (there is no CPU fallback)"
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        if os.environ.get("QASR_LIB_PATH") and not hasattr(lib, name):
            continue  # A/B build of an older revision: tolerate symbols it does not have yet
        fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error(handle=None) -> str:
    msg = load().qasr_last_error(handle)
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, handle=None, decoder: bool = False) -> None:
    """Map a qasr_status to the exception type the reference raises for the same condition."""
    if rc == QASR_OK:
        return
    if decoder:
        raw = load().qasr_decoder_last_error(handle)
        msg = (raw.decode("utf-8", "replace") if raw else "") or f"libqasr error {rc}"
    else:
        msg = last_error(handle) or f"libqasr error {rc}"
    if rc == QASR_ERR_INVALID:
        raise ValueError(msg)
    if rc == QASR_ERR_NOMEM:
        raise MemoryError(msg)
    raise QasrError(f"[{rc}] {msg}")
