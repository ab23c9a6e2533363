"""Audio encoder — B200 path behind the reference's ``encoder`` module interface
(src/qwen3_asr_mlx/encoder.py).

``AudioEncoder(config)(mel)`` keeps the reference contract: a ``(n_mels, T)`` (or
``(1, n_mels, T)``) log-mel in, ``(1, n_tokens, output_dim)`` projected embeddings out.
The forward runs entirely inside libqasr (conv stem as implicit tcgen05 GEMM, 24 windowed-
attention transformer layers, projector).  ``encode_batch`` / ``encode_audio_batch`` are the
varlen-packed batched additions; their result equals a per-utterance loop of ``__call__``.
"""
from __future__ import annotations

import ctypes
from pathlib import Path
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, runtime, weights as _weights
from ._array import DeviceArray, as_device_f32
from .audio import HOP_LENGTH, N_MELS, _as_waveform, SAMPLE_RATE, pack_waveforms
from .config import AudioEncoderConfig


class SinusoidalPositionEmbedding:
    """Sinusoidal table; ``__call__(seqlen)`` returns its first rows (reference encoder.py:21-44).

    The table is produced by the library (``qasr_positional_embedding``) so that tests see the
    values the kernels add.
    """

    def __init__(self, max_positions: int, d_model: int, _handle: Optional[runtime.Handle] = None):
        self.max_positions = max_positions
        self.d_model = d_model
        self._handle = _handle
        self._table: Optional[np.ndarray] = None

    def _ensure(self) -> np.ndarray:
        if self._table is None:
            h = self._handle or runtime.Handle(AudioEncoderConfig(d_model=self.d_model, encoder_attention_heads=self.d_model // 64,
                                                                  encoder_layers=0, max_source_positions=self.max_positions))
            out = np.empty((self.max_positions, self.d_model), dtype=np.float32)
            h.check(h.lib.qasr_positional_embedding(h.ptr, self.max_positions, out.ctypes.data_as(ctypes.POINTER(ctypes.c_float))))
            self._table = out
        return self._table

    def __call__(self, seqlen: int) -> np.ndarray:
        return self._ensure()[:seqlen, :]


class AudioEncoder:
    """Qwen3-ASR audio encoder on one B200 (reference encoder.py:129-323).

    Parameters
    ----------
    config: ``AudioEncoderConfig`` (defaults = 1.7B model).
    seed:   seed of the random initialisation used until weights are loaded (the reference's
            freshly constructed module is randomly initialised by MLX).
    device: CUDA device index (default: LOCAL_RANK or the current device).
    """

    def __init__(self, config: AudioEncoderConfig, seed: int = 0, device: Optional[int] = None):
        self.config = config
        self.chunk_size = config.n_window * 2
        self._seed = seed
        self._handle = runtime.Handle(config, device)
        self._ready = False
        self.positional_embedding = SinusoidalPositionEmbedding(config.max_source_positions, config.d_model, self._handle)

    # ------------------------------------------------------------------ helpers kept from the reference
    @staticmethod
    def _conv_output_length(input_length: int) -> int:
        """Length after three k3/s2/p1 convolutions: L -> (L - 1)//2 + 1, thrice (encoder.py:197-207)."""
        n = input_length
        for _ in range(3):
            n = (n - 1) // 2 + 1
        return n

    @staticmethod
    def _block_attention_mask(seq_len: int, cu_seqlens: List[int]) -> Optional[np.ndarray]:
        """Dense additive block-diagonal mask ``(1, 1, n, n)`` or None for a single block
        (encoder.py:209-229).  The kernels never build this; it exists for interface parity."""
        if len(cu_seqlens) <= 2:
            return None
        mask = np.full((seq_len, seq_len), -1e9, dtype=np.float32)
        for lo, hi in zip(cu_seqlens[:-1], cu_seqlens[1:]):
            mask[lo:hi, lo:hi] = 0.0
        return mask[None, None]

    def num_tokens(self, n_frames: int) -> int:
        full, rem = divmod(int(n_frames), self.chunk_size)
        return full * self._conv_output_length(self.chunk_size) + (self._conv_output_length(rem) if rem else 0)

    # ------------------------------------------------------------------ weights
    def load_weights(self, items: Iterable[Tuple[str, np.ndarray]] | Dict[str, np.ndarray]) -> None:
        """Install parameters (reference names, reference layouts) and finalise the device copies."""
        if self._ready:
            raise _lib.QasrError("weights already loaded for this encoder; construct a new AudioEncoder")
        h = self._handle
        pairs = items.items() if isinstance(items, dict) else items
        for name, value in pairs:
            if isinstance(value, torch.Tensor):
                value = value.detach().float().cpu().numpy()
            arr = np.ascontiguousarray(np.asarray(value), dtype=np.float32)
            shape = (ctypes.c_int64 * arr.ndim)(*arr.shape)
            h.check(h.lib.qasr_set_weight(h.ptr, name.encode(), arr.ctypes.data_as(ctypes.c_void_p), _lib.QASR_F32, arr.ndim, shape))
        h.check(h.lib.qasr_finalize_weights(h.ptr))
        self._ready = True

    def _ensure_weights(self) -> None:
        if not self._ready:
            self.load_weights(_weights.random_init(self.config, seed=self._seed))

    # ------------------------------------------------------------------ forward
    def encode_batch(self, mels: Sequence, out_dtype: str = "float32") -> Tuple[DeviceArray, np.ndarray]:
        """Encode a batch of ``(128, T_u)`` log-mels, varlen-packed.

        Returns ``(embeddings (sum n_u, output_dim), token_offsets (B+1,))``.
        """
        self._ensure_weights()
        h = self._handle
        tens = []
        for m in mels:
            t = as_device_f32(m, h.torch_device)
            if t.ndim == 3:
                t = t[0]  # the reference drops batch entries > 0 (encoder.py:249-250)
            if t.ndim != 2 or t.shape[0] != self.config.num_mel_bins:
                raise ValueError(f"mel must have shape ({self.config.num_mel_bins}, T), got {tuple(t.shape)}")
            tens.append(t.contiguous())
        if not tens:
            raise ValueError("empty batch")
        foffs = runtime.offsets_array([int(t.shape[1]) for t in tens])
        with torch.cuda.device(h.torch_device):
            packed = tens[0].reshape(-1) if len(tens) == 1 else torch.cat([t.reshape(-1) for t in tens])
            return self._encode_packed(packed, foffs, out_dtype)

    def _encode_packed(self, packed_mel: torch.Tensor, foffs: np.ndarray, out_dtype: str) -> Tuple[DeviceArray, np.ndarray]:
        h = self._handle
        B = len(foffs) - 1
        n_tok = sum(self.num_tokens(int(foffs[u + 1] - foffs[u])) for u in range(B))
        tdt, cdt = (torch.bfloat16, _lib.QASR_BF16) if out_dtype in ("bfloat16", "bf16") else (torch.float32, _lib.QASR_F32)
        out = torch.empty((n_tok, self.config.output_dim), dtype=tdt, device=h.torch_device)
        toffs = np.zeros(B + 1, dtype=np.int64)
        h.check(h.lib.qasr_encode(h.ptr, ctypes.c_void_p(packed_mel.data_ptr()), runtime.i64_ptr(foffs), B,
                                  ctypes.c_void_p(out.data_ptr()), cdt, runtime.i64_ptr(toffs), h.stream_ptr()))
        return DeviceArray(out), toffs

    def encode_audio_batch(self, audios: Sequence, out_dtype: str = "float32", max_tokens_per_call: int = 65536) -> Tuple[DeviceArray, np.ndarray]:
        """Waveforms in, packed embeddings out: mel + encoder back to back on the device
        (the reference call site model.py:331-335, batched).  Batches larger than ``max_tokens_per_call``
        audio tokens are processed in consecutive sub-batches so the activation workspace stays bounded."""
        self._ensure_weights()
        h = self._handle
        waves = [_as_waveform(a, SAMPLE_RATE) for a in audios]
        if not waves:
            raise ValueError("empty batch")
        lengths = [int(w.shape[0]) for w in waves]
        for n in lengths:
            if n < HOP_LENGTH:
                raise ValueError(f"zero-size array to reduction operation maximum which has no identity (audio of {n} samples < {HOP_LENGTH})")
        tokens = [self.num_tokens(n // HOP_LENGTH) for n in lengths]
        if sum(tokens) > max_tokens_per_call and len(waves) > 1:
            pieces, offsets, cur, acc = [], [0], [], 0
            for w, t in zip(waves, tokens):
                if cur and acc + t > max_tokens_per_call:
                    emb, toffs = self.encode_audio_batch(cur, out_dtype, max_tokens_per_call=1 << 62)
                    pieces.append(emb.tensor)
                    base = offsets[-1]
                    offsets.extend(base + int(v) for v in toffs[1:])
                    cur, acc = [], 0
                cur.append(w)
                acc += t
            emb, toffs = self.encode_audio_batch(cur, out_dtype, max_tokens_per_call=1 << 62)
            pieces.append(emb.tensor)
            base = offsets[-1]
            offsets.extend(base + int(v) for v in toffs[1:])
            return DeviceArray(torch.cat(pieces)), np.asarray(offsets, dtype=np.int64)
        with torch.cuda.device(h.torch_device):
            packed, soffs = pack_waveforms(h, waves)
            return self.encode_packed_audio(packed, soffs, out_dtype)

    def encode_packed_audio(self, packed_audio: torch.Tensor, soffs: np.ndarray, out_dtype: str = "float32",
                            out: Optional[torch.Tensor] = None) -> Tuple[DeviceArray, np.ndarray]:
        """Device-resident packed audio (float32, ``soffs`` sample offsets) -> packed embeddings.

        Passing the same ``packed_audio`` / ``out`` buffers and offsets again replays the CUDA graph
        libqasr captured for that call instead of re-launching ~230 kernels."""
        self._ensure_weights()
        h = self._handle
        B = len(soffs) - 1
        n_tok = sum(self.num_tokens(int(soffs[u + 1] - soffs[u]) // HOP_LENGTH) for u in range(B))
        tdt, cdt = (torch.bfloat16, _lib.QASR_BF16) if out_dtype in ("bfloat16", "bf16") else (torch.float32, _lib.QASR_F32)
        if out is None:
            out = torch.empty((n_tok, self.config.output_dim), dtype=tdt, device=h.torch_device)
        elif tuple(out.shape) != (n_tok, self.config.output_dim) or out.dtype != tdt or not out.is_contiguous() or out.data_ptr() % 16:
            raise ValueError(f"out must be a contiguous, 16-byte aligned {tdt} tensor of shape {(n_tok, self.config.output_dim)}")
        toffs = np.zeros(B + 1, dtype=np.int64)
        h.check(h.lib.qasr_encode_audio(h.ptr, ctypes.c_void_p(packed_audio.data_ptr()), runtime.i64_ptr(soffs), B,
                                        ctypes.c_void_p(out.data_ptr()), cdt, runtime.i64_ptr(toffs), h.stream_ptr()))
        return DeviceArray(out), toffs

    def encode_packed_audio_hidden(self, packed_audio: torch.Tensor, soffs: np.ndarray) -> Tuple[int, np.ndarray]:
        """``qasr_encode_audio_hidden``: mel + conv stem + transformer layers of a packed batch (encoder.py:235-317); the
        final hidden states stay in the handle for :meth:`project_rows`.  Returns ``(n_tokens, token_offsets)``."""
        self._ensure_weights()
        h = self._handle
        B = len(soffs) - 1
        toffs = np.zeros(B + 1, dtype=np.int64)
        h.check(h.lib.qasr_encode_audio_hidden(h.ptr, ctypes.c_void_p(packed_audio.data_ptr()), runtime.i64_ptr(soffs), B,
                                               runtime.i64_ptr(toffs), h.stream_ptr()))
        return int(toffs[-1]), toffs

    def project_rows(self, row0: int, out: torch.Tensor) -> torch.Tensor:
        """``qasr_project_rows``: ln_post -> proj1 -> GELU -> proj2 (encoder.py:319-321) for rows
        ``[row0, row0 + len(out))`` of the last :meth:`encode_packed_audio_hidden` call, written into ``out``
        (contiguous ``(rows, output_dim)`` float32 / bfloat16 CUDA tensor).  Any blocking of the rows gives the same bits
        as one :meth:`encode_packed_audio` call."""
        h = self._handle
        if out.ndim != 2 or out.shape[1] != self.config.output_dim or not out.is_contiguous() or out.data_ptr() % 16 or \
                out.dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("out must be a contiguous, 16-byte aligned (rows, output_dim) float32 / bfloat16 tensor")
        cdt = _lib.QASR_BF16 if out.dtype == torch.bfloat16 else _lib.QASR_F32
        h.check(h.lib.qasr_project_rows(h.ptr, int(row0), int(out.shape[0]), ctypes.c_void_p(out.data_ptr()), cdt, h.stream_ptr()))
        return out

    def find_split_points(self, audio: torch.Tensor, chunk_samples: int, search_samples: int, frame_samples: int = 480,
                          return_energy: bool = False):
        """Cut positions for long audio on the device (reference ``_find_split_points``, model.py:454-513):
        per-frame float32 RMS energy (bit-identical to the reference's numpy expression) and the first
        lowest-energy frame within +-``search_samples`` of every multiple of ``chunk_samples``.
        ``audio`` is a 1-D float32 CUDA tensor; returns a list of ints (and the energies if asked)."""
        h = self._handle
        if not isinstance(audio, torch.Tensor) or not audio.is_cuda or audio.dtype != torch.float32 or audio.ndim != 1 or not audio.is_contiguous():
            raise ValueError("audio must be a contiguous 1-D float32 CUDA tensor")
        if chunk_samples <= 0 or search_samples < 0 or frame_samples <= 0:
            raise ValueError("chunk_samples and frame_samples must be positive")
        n = int(audio.shape[0])
        if n // frame_samples == 0:  # no whole frame: the reference returns [] (model.py:486-487)
            return ([], torch.empty(0, dtype=torch.float32, device=audio.device)) if return_energy else []
        max_points = max(1, (n - 1) // int(chunk_samples))
        points = np.zeros(max_points, dtype=np.int64)
        count = ctypes.c_int32(0)
        energy = torch.empty(n // frame_samples, dtype=torch.float32, device=audio.device) if return_energy else None
        h.check(h.lib.qasr_find_split_points(h.ptr, ctypes.c_void_p(audio.data_ptr()), n, int(chunk_samples), int(search_samples),
                                             int(frame_samples), runtime.i64_ptr(points), max_points, ctypes.byref(count),
                                             ctypes.c_void_p(energy.data_ptr()) if energy is not None and energy.numel() else None,
                                             h.stream_ptr()))
        pts = [int(v) for v in points[: count.value]]
        return (pts, energy) if return_energy else pts

    def encode_audio_host(self, audio: np.ndarray, soffs: np.ndarray, out: np.ndarray) -> np.ndarray:
        """Host buffers in and out through ``qasr_encode_audio_host`` (H2D and D2H inside the call)."""
        self._ensure_weights()
        h = self._handle
        B = len(soffs) - 1
        toffs = np.zeros(B + 1, dtype=np.int64)
        cdt = _lib.QASR_F32 if out.dtype == np.float32 else _lib.QASR_BF16
        h.check(h.lib.qasr_encode_audio_host(h.ptr, ctypes.c_void_p(audio.ctypes.data), runtime.i64_ptr(soffs), B,
                                             ctypes.c_void_p(out.ctypes.data), cdt, runtime.i64_ptr(toffs)))
        return toffs

    def encode_audio_host_async(self, slot: int, audio: np.ndarray, soffs: np.ndarray, out: np.ndarray) -> np.ndarray:
        """Submit one batch into pipeline slot 0/1 (``qasr_encode_audio_host_async``): its H2D / kernels /
        D2H overlap the other slot's.  ``out`` is valid after ``host_wait(slot)``; returns token offsets."""
        self._ensure_weights()
        h = self._handle
        toffs = np.zeros(len(soffs), dtype=np.int64)
        cdt = _lib.QASR_F32 if out.dtype == np.float32 else _lib.QASR_BF16
        h.check(h.lib.qasr_encode_audio_host_async(h.ptr, int(slot), ctypes.c_void_p(audio.ctypes.data), runtime.i64_ptr(soffs),
                                                   len(soffs) - 1, ctypes.c_void_p(out.ctypes.data), cdt, runtime.i64_ptr(toffs)))
        return toffs

    def host_wait(self, slot: int) -> None:
        self._handle.check(self._handle.lib.qasr_host_wait(self._handle.ptr, int(slot)))

    def __call__(self, mel) -> DeviceArray:
        """``(n_mels, T)`` or ``(batch, n_mels, T)`` log-mel -> ``(1, n_tokens, output_dim)``."""
        emb, _ = self.encode_batch([mel])
        return DeviceArray(emb.tensor.unsqueeze(0))

    # ------------------------------------------------------------------ test hooks / lifecycle
    def set_debug(self, enabled: bool) -> None:
        self._handle.check(self._handle.lib.qasr_set_debug(self._handle.ptr, int(enabled)))

    def debug_read(self, what: str, n_tokens: int) -> np.ndarray:
        out = np.empty((n_tokens, self.config.d_model), dtype=np.float32)
        self._handle.check(self._handle.lib.qasr_debug_read(self._handle.ptr, what.encode(),
                                                            out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), out.size))
        return out

    def stats(self) -> Dict[str, int]:
        return self._handle.stats()

    def set_profile(self, enabled: bool) -> None:
        """Bracket every kernel launch with CUDA events on the launch stream (resets the totals)."""
        self._handle.check(self._handle.lib.qasr_set_profile(self._handle.ptr, int(enabled)))

    def get_profile(self) -> Dict[str, Dict[str, float]]:
        """Per-kernel-category totals since ``set_profile``: device ms, algorithmic flops/bytes, launches."""
        prof = _lib.QasrProfile()
        self._handle.check(self._handle.lib.qasr_get_profile(self._handle.ptr, ctypes.byref(prof)))
        out = {}
        for i in range(_lib.PROF_CATEGORIES):
            name = self._handle.lib.qasr_profile_name(i).decode()
            out[name] = {"ms": prof.ms[i], "flops": prof.flops[i], "bytes": prof.bytes[i], "launches": int(prof.launches[i])}
        return out

    def reserve(self, total_frames: int, batch: int) -> None:
        self._handle.check(self._handle.lib.qasr_reserve(self._handle.ptr, int(total_frames), int(batch)))

    def close(self) -> None:
        self._handle.close()


def load_encoder_weights(model: AudioEncoder, model_path) -> None:
    """Populate ``model`` from ``<model_path>/model.safetensors`` (reference encoder.py:330-359):
    keys with prefix ``audio_tower.`` are kept and the prefix stripped; layouts are used as stored
    (Conv2d weights (O, kH, kW, I)).  ``model_path`` is a local directory or a hub repo id (``snapshot_download``,
    encoder.py:342-344)."""
    from ._hub import model_dir

    path = model_dir(model_path)
    model.load_weights(_weights.load_safetensors(path / "model.safetensors"))
