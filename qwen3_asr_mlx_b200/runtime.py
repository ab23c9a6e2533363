"""Per-process runtime: one libqasr handle per (device, purpose).

The library is the only compute backend.  torch is used here for device buffers, streams and the
current-device bookkeeping (``torch.cuda.current_stream()`` is what kernels are launched on).
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .config import AudioEncoderConfig


def require_cuda() -> None:
    if not torch.cuda.is_available():
        raise _lib.QasrError("no CUDA device visible: the B200 path has no CPU fallback")


def local_device() -> int:
    """Device index for this process: LOCAL_RANK when launched by torchrun, else the current device."""
    require_cuda()
    if "LOCAL_RANK" in os.environ:
        return int(os.environ["LOCAL_RANK"]) % torch.cuda.device_count()
    return torch.cuda.current_device()


def to_c_config(cfg: AudioEncoderConfig) -> _lib.QasrConfig:
    return _lib.QasrConfig(
        cfg.d_model, cfg.encoder_layers, cfg.encoder_attention_heads, cfg.encoder_ffn_dim, cfg.num_mel_bins,
        cfg.max_source_positions, cfg.output_dim, cfg.n_window, cfg.n_window_infer, cfg.downsample_hidden_size,
    )


def offsets_array(lengths: Sequence[int]) -> np.ndarray:
    out = np.zeros(len(lengths) + 1, dtype=np.int64)
    np.cumsum(np.asarray(lengths, dtype=np.int64), out=out[1:])
    return out


def i64_ptr(a: np.ndarray):
    assert a.dtype == np.int64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))


class Handle:
    """RAII wrapper of ``qasr_handle*``."""

    def __init__(self, cfg: AudioEncoderConfig, device: Optional[int] = None):
        require_cuda()
        self.lib = _lib.load()
        self.device = local_device() if device is None else int(device)
        self.cfg = cfg
        self._h = ctypes.c_void_p()
        c = to_c_config(cfg)
        _lib.check(self.lib.qasr_create(self.device, ctypes.byref(c), ctypes.byref(self._h)))
        self.torch_device = torch.device("cuda", self.device)

    @property
    def ptr(self) -> ctypes.c_void_p:
        if not self._h:
            raise _lib.QasrError("handle already destroyed")
        return self._h

    def check(self, rc: int) -> None:
        _lib.check(rc, self._h)

    def stream_ptr(self) -> ctypes.c_void_p:
        return ctypes.c_void_p(torch.cuda.current_stream(self.torch_device).cuda_stream)

    def stats(self) -> Dict[str, int]:
        s = _lib.QasrStats()
        self.check(self.lib.qasr_get_stats(self.ptr, ctypes.byref(s)))
        return {"kernel_launches": s.kernel_launches, "workspace_bytes": s.workspace_bytes, "weight_bytes": s.weight_bytes}

    def close(self) -> None:
        if getattr(self, "_h", None):
            self.lib.qasr_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_frontend: Dict[int, Handle] = {}


def frontend_handle(device: Optional[int] = None) -> Handle:
    """Weight-less handle used by the mel frontend (tables only), cached per device."""
    dev = local_device() if device is None else int(device)
    h = _frontend.get(dev)
    if h is None:
        h = Handle(AudioEncoderConfig(encoder_layers=0), dev)
        _frontend[dev] = h
    return h


def release_all() -> None:
    for h in _frontend.values():
        h.close()
    _frontend.clear()
