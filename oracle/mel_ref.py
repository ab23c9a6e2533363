"""The reference's own mel frontend, executed verbatim (authoring container only).

Loads /root/reference/src/qwen3_asr_mlx/audio.py unmodified behind the `mlx.core` stand-in in
oracle/_mlx_shim.  /root/reference does not exist on the GPU box, so nothing at run time there
may import this module; it is used by oracle/gen_golden.py and by the CPU tests that pin
oracle/mel_np.py (skipped when the reference tree is absent).
"""
from __future__ import annotations

import importlib.util
import os
import sys

REFERENCE_AUDIO = "/root/reference/src/qwen3_asr_mlx/audio.py"
_mod = None


def available() -> bool:
    return os.path.exists(REFERENCE_AUDIO)


def module():
    global _mod
    if _mod is None:
        shim = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_mlx_shim")
        if "mlx" not in sys.modules:
            sys.path.insert(0, shim)
        spec = importlib.util.spec_from_file_location("_qwen3_asr_mlx_reference_audio", REFERENCE_AUDIO)
        _mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_mod)
    return _mod


def log_mel_spectrogram(audio):
    return module().log_mel_spectrogram(audio)


def mel_filterbank():
    return module()._get_mel_filterbank()
