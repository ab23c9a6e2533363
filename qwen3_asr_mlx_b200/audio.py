"""Mel-spectrogram frontend — B200 path behind the reference's ``audio`` module interface
(src/qwen3_asr_mlx/audio.py).

``log_mel_spectrogram`` keeps the reference signature (audio.py:238-246) but runs the
framing / Hann / 400-point FFT / power / mel / log10 / max-8 clamp pipeline as CUDA kernels
(csrc/mel.cuh) through ``qasr_mel``.  ``log_mel_spectrogram_batch`` is the batched addition.
The numpy helpers below (filterbank construction, WAV decode) are host-side table / file
utilities kept for interface parity; they are not on the compute path.
"""
from __future__ import annotations

import ctypes
import struct
from pathlib import Path
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import runtime
from ._array import DeviceArray, as_device_f32

# Qwen3-ASR frontend constants (reference audio.py:15-21)
SAMPLE_RATE = 16_000
N_FFT = 400
HOP_LENGTH = 160
N_MELS = 128
F_MIN = 0.0
F_MAX = 8_000.0

_mel_filterbank_cache: dict[tuple, np.ndarray] = {}


# --------------------------------------------------------------------------- filterbank (host table)
def _hz_to_mel(freq):
    """HTK mel scale, as the reference computes it (audio.py:31-33)."""
    return 2595.0 * np.log10(1.0 + freq / 700.0)


def _mel_to_hz(mel):
    """Inverse of ``_hz_to_mel`` (audio.py:36-38)."""
    return 700.0 * (10.0 ** (mel / 2595.0) - 1.0)


def _build_mel_filterbank(n_fft: int, n_mels: int, sample_rate: int, f_min: float, f_max: float) -> np.ndarray:
    """(n_mels, n_fft//2+1) float32 triangular filterbank, each row divided by its width in Hz.

    Same construction as the reference (audio.py:41-80), vectorised over the filters: the
    slopes are evaluated in float64, stored as float32, then divided (in float64) by
    ``f_right - f_left`` and rounded to float32 again.
    """
    freqs = np.linspace(0.0, sample_rate / 2.0, n_fft // 2 + 1)
    edges = _mel_to_hz(np.linspace(_hz_to_mel(f_min), _hz_to_mel(f_max), n_mels + 2))
    left, centre, right = edges[:-2, None], edges[1:-1, None], edges[2:, None]
    rising = (freqs[None, :] - left) / (centre - left)
    falling = (right - freqs[None, :]) / (right - centre)
    tri = np.maximum(0.0, np.minimum(rising, falling)).astype(np.float32)
    width = (right - left)[:, 0]
    scaled = np.where(width[:, None] > 0.0, tri.astype(np.float64) / np.where(width > 0.0, width, 1.0)[:, None], tri)
    return scaled.astype(np.float32)


def _get_mel_filterbank(n_fft: int = N_FFT, n_mels: int = N_MELS, sample_rate: int = SAMPLE_RATE,
                        f_min: float = F_MIN, f_max: float = F_MAX) -> np.ndarray:
    """Cached filterbank (reference audio.py:83-96)."""
    key = (n_fft, n_mels, sample_rate, f_min, f_max)
    fb = _mel_filterbank_cache.get(key)
    if fb is None:
        fb = _mel_filterbank_cache[key] = _build_mel_filterbank(n_fft, n_mels, sample_rate, f_min, f_max)
    return fb


# --------------------------------------------------------------------------- file decode (host utility)
def _read_wav_pcm(path) -> Tuple[np.ndarray, int]:
    """Decode a RIFF/WAVE file holding PCM16, PCM32 or IEEE float32 samples to mono float32.

    Covers the same formats as the reference's fast path (audio.py:103-170); anything else raises
    ValueError.  Multi-channel audio is averaged to mono.
    """
    data = Path(path).read_bytes()
    if len(data) < 12 or data[0:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError(f"{path}: not a RIFF/WAVE file")
    pos, fmt, payload = 12, None, None
    while pos + 8 <= len(data):
        tag, size = data[pos:pos + 4], struct.unpack_from("<I", data, pos + 4)[0]
        body = data[pos + 8: pos + 8 + size]
        if tag == b"fmt ":
            fmt = struct.unpack_from("<HHIIHH", body, 0)
        elif tag == b"data":
            payload = body
        pos += 8 + size + (size & 1)
    if fmt is None or payload is None:
        raise ValueError(f"{path}: missing fmt or data chunk")
    code, channels, rate, _, _, bits = fmt
    if code == 1 and bits == 16:
        x = np.frombuffer(payload, dtype="<i2").astype(np.float32) / 32768.0
    elif code == 1 and bits == 32:
        x = np.frombuffer(payload, dtype="<i4").astype(np.float32) / 2147483648.0
    elif code == 3 and bits == 32:
        x = np.frombuffer(payload, dtype="<f4").astype(np.float32)
    else:
        raise ValueError(f"{path}: unsupported WAV encoding (format {code}, {bits}-bit)")
    if channels > 1:
        n = len(x) // channels
        x = x[: n * channels].reshape(n, channels).mean(axis=1)
    return x.astype(np.float32), int(rate)


def load_audio(path, target_sr: int = SAMPLE_RATE) -> np.ndarray:
    """Load an audio file as mono float32 at ``target_sr`` (reference audio.py:173-204).

    WAV is decoded natively; other containers need ``soundfile`` (imported lazily, as in the
    reference).  Sample-rate conversion is the reference's linear interpolation.
    """
    path = Path(path)
    samples: Optional[np.ndarray] = None
    rate: Optional[int] = None
    if path.suffix.lower() == ".wav":
        try:
            samples, rate = _read_wav_pcm(path)
        except Exception:  # any failure of the fast path falls through to soundfile, as in the reference (audio.py:189-193)
            samples = None
    if samples is None:
        import soundfile as sf  # optional dependency, exactly like the reference

        samples, rate = sf.read(str(path), dtype="float32", always_2d=False)
        if samples.ndim == 2:
            samples = samples.mean(axis=1)
    if rate != target_sr:
        n_out = int(len(samples) * target_sr / rate)
        src_pos = np.linspace(0.0, len(samples) - 1, n_out)
        samples = np.interp(src_pos, np.arange(len(samples)), samples).astype(np.float32)
    return samples


# --------------------------------------------------------------------------- the hot path
def _check_frontend_params(n_fft, hop_length, n_mels, sample_rate, f_min, f_max) -> None:
    if (n_fft, hop_length, n_mels, sample_rate, float(f_min), float(f_max)) != (N_FFT, HOP_LENGTH, N_MELS, SAMPLE_RATE, F_MIN, F_MAX):
        raise ValueError(
            "the B200 mel kernels implement the Qwen3-ASR frontend only "
            f"(n_fft={N_FFT}, hop_length={HOP_LENGTH}, n_mels={N_MELS}, sample_rate={SAMPLE_RATE}, f_min={F_MIN}, f_max={F_MAX})"
        )


def _as_waveform(audio, sample_rate: int) -> np.ndarray | torch.Tensor:
    if isinstance(audio, (str, Path)):
        audio = load_audio(audio, target_sr=sample_rate)
    if isinstance(audio, DeviceArray):
        audio = audio.tensor
    if isinstance(audio, torch.Tensor):
        if audio.ndim != 1:
            raise ValueError(f"Audio array must be 1-D (mono), got shape {tuple(audio.shape)}")
        return audio
    a = np.asarray(audio, dtype=np.float32)
    if a.ndim != 1:
        raise ValueError(f"Audio array must be 1-D (mono), got shape {a.shape}")
    return a


def pack_waveforms(h, waves: Sequence) -> Tuple[torch.Tensor, np.ndarray]:
    """Varlen-pack a batch of 1-D waveforms into ONE contiguous float32 device buffer + sample offsets, without a per-utterance
    host loop of device copies: host arrays are concatenated on the host and uploaded with one H2D copy; device tensors that
    already lie back to back in one allocation are used in place (zero copy); any other set of device tensors goes through
    ``qasr_pack_audio`` (one pointer-table upload + one kernel, whatever the batch size)."""
    soffs = np.zeros(len(waves) + 1, dtype=np.int64)
    np.cumsum(np.fromiter((w.shape[0] for w in waves), dtype=np.int64, count=len(waves)), out=soffs[1:])
    dev = h.torch_device
    if len(waves) == 1:
        return as_device_f32(waves[0], dev), soffs
    if all(isinstance(w, np.ndarray) for w in waves):
        host = np.concatenate([np.asarray(w, dtype=np.float32) for w in waves])
        return torch.from_numpy(host).to(dev, non_blocking=True), soffs
    tens = [w if (isinstance(w, torch.Tensor) and w.device == dev and w.dtype == torch.float32 and w.is_contiguous())
            else as_device_f32(w, dev) for w in waves]
    base = tens[0]
    addr = np.fromiter((t.data_ptr() for t in tens), dtype=np.int64, count=len(tens))
    # back to back AND inside the first tensor's allocation (no other allocation can lie inside its address range)
    st = base.untyped_storage()
    adjacent = bool(np.array_equal(addr[1:], addr[:-1] + 4 * soffs[1:-1] - 4 * soffs[:-2])) and \
        st.data_ptr() <= int(addr[0]) and int(addr[0]) + 4 * int(soffs[-1]) <= st.data_ptr() + st.nbytes()
    if adjacent:
        return torch.as_strided(base, (int(soffs[-1]),), (1,)), soffs
    packed = torch.empty(int(soffs[-1]), dtype=torch.float32, device=dev)
    ptrs = (ctypes.c_void_p * len(tens)).from_buffer(addr.astype(np.uint64))
    h.check(h.lib.qasr_pack_audio(h.ptr, ptrs, runtime.i64_ptr(soffs), len(tens), ctypes.c_void_p(packed.data_ptr()), h.stream_ptr()))
    # the sources must stay alive until the kernel has read them: they are referenced by `tens` until this frame returns and the
    # caching allocator only hands their memory to later work on the same stream
    return packed, soffs


def log_mel_spectrogram_batch(audios: Sequence, device: Optional[int] = None) -> Tuple[DeviceArray, np.ndarray]:
    """Log-mel features of a batch of utterances in one launch.

    Returns ``(mel, frame_offsets)``: ``mel`` is a flat float32 device array in which utterance
    ``u`` is the row-major ``(128, T_u)`` block starting at ``128 * frame_offsets[u]``; each
    block equals the reference's ``log_mel_spectrogram(audios[u])``.
    """
    h = runtime.frontend_handle(device)
    waves = [_as_waveform(a, SAMPLE_RATE) for a in audios]
    if not waves:
        raise ValueError("empty batch")
    lengths = np.fromiter((w.shape[0] for w in waves), dtype=np.int64, count=len(waves))
    if int(lengths.min()) < HOP_LENGTH:
        # the reference fails here with numpy's "zero-size array to reduction operation maximum"
        raise ValueError(f"zero-size array to reduction operation maximum which has no identity (audio of {int(lengths.min())} samples < {HOP_LENGTH})")
    foffs = np.zeros(len(waves) + 1, dtype=np.int64)
    np.cumsum(lengths // HOP_LENGTH, out=foffs[1:])
    with torch.cuda.device(h.torch_device):
        packed, soffs = pack_waveforms(h, waves)
        mel = torch.empty(int(foffs[-1]) * N_MELS, dtype=torch.float32, device=h.torch_device)
        h.check(h.lib.qasr_mel(h.ptr, ctypes.c_void_p(packed.data_ptr()), runtime.i64_ptr(soffs), len(waves),
                               ctypes.c_void_p(mel.data_ptr()), h.stream_ptr()))
    return DeviceArray(mel), foffs


def log_mel_spectrogram_packed(packed_audio: torch.Tensor, sample_offsets) -> Tuple[DeviceArray, np.ndarray]:
    """Log-mel features of a varlen-packed batch that is already resident on the device: ``packed_audio`` is a contiguous 1-D
    float32 CUDA tensor, utterance ``u`` = ``[sample_offsets[u], sample_offsets[u+1])`` (the layout ``qasr_mel`` takes).  No
    per-utterance host work beyond the offset arithmetic inside the library; returns ``(mel, frame_offsets)`` as
    ``log_mel_spectrogram_batch`` does."""
    if not isinstance(packed_audio, torch.Tensor) or not packed_audio.is_cuda or packed_audio.dtype != torch.float32 \
            or packed_audio.ndim != 1 or not packed_audio.is_contiguous():
        raise ValueError("packed_audio must be a contiguous 1-D float32 CUDA tensor")
    soffs = np.ascontiguousarray(np.asarray(sample_offsets, dtype=np.int64))
    if soffs.ndim != 1 or len(soffs) < 2 or soffs[0] != 0 or int(soffs[-1]) != packed_audio.numel():
        raise ValueError("sample_offsets must start at 0 and end at the number of samples")
    lengths = np.diff(soffs)
    if (lengths < HOP_LENGTH).any():
        raise ValueError(f"zero-size array to reduction operation maximum which has no identity (an utterance has < {HOP_LENGTH} samples)")
    h = runtime.frontend_handle(packed_audio.device.index)
    foffs = np.zeros(len(soffs), dtype=np.int64)
    np.cumsum(lengths // HOP_LENGTH, out=foffs[1:])
    with torch.cuda.device(h.torch_device):
        mel = torch.empty(int(foffs[-1]) * N_MELS, dtype=torch.float32, device=h.torch_device)
        h.check(h.lib.qasr_mel(h.ptr, ctypes.c_void_p(packed_audio.data_ptr()), runtime.i64_ptr(soffs), len(soffs) - 1,
                               ctypes.c_void_p(mel.data_ptr()), h.stream_ptr()))
    return DeviceArray(mel), foffs


def log_mel_spectrogram(audio, n_fft: int = N_FFT, hop_length: int = HOP_LENGTH, n_mels: int = N_MELS,
                        sample_rate: int = SAMPLE_RATE, f_min: float = F_MIN, f_max: float = F_MAX) -> DeviceArray:
    """Log-mel spectrogram of one utterance, shape ``(n_mels, n_samples // 160)`` float32 on the GPU.

    Same signature and result as the reference (audio.py:238-278): STFT (n_fft 400, hop 160,
    symmetric Hann, reflect padding), power with the last frame dropped, mel filterbank,
    ``log10(max(., 1e-10))``, clamp to ``max - 8`` over the whole utterance, ``(x + 4) / 4``.
    """
    _check_frontend_params(n_fft, hop_length, n_mels, sample_rate, f_min, f_max)
    mel, foffs = log_mel_spectrogram_batch([audio])
    return DeviceArray(mel.tensor.view(N_MELS, int(foffs[1])))
