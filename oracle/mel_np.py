"""numpy restatement of the reference mel frontend (TEST INFRASTRUCTURE — see oracle/__init__.py).

Follows /root/reference/src/qwen3_asr_mlx/audio.py:
  hz/mel maps            :31-38     HTK formula (the docstrings say Slaney)
  filterbank             :41-80     triangles on linspace(0, 8000, 201), float32, / width in Hz
  stft                   :211-235   symmetric Hann (np.hanning), reflect pad 200, hop 160, rfft n=400
  log_mel_spectrogram    :238-278   |.|^2 with last frame dropped, fb @ power, log10(max(.,1e-10)),
                                    max(., global max - 8), (. + 4) / 4
`log_mel_spectrogram` keeps the reference's per-frame loop (bit-exact with the reference under the
same numpy); `log_mel_spectrogram_fast` batches the FFT (same pocketfft kernels, one call).
"""
from __future__ import annotations

import numpy as np

SAMPLE_RATE, N_FFT, HOP, N_MELS = 16000, 400, 160, 128


def hz_to_mel(f):
    return 2595.0 * np.log10(1.0 + f / 700.0)  # audio.py:31-33


def mel_to_hz(m):
    return 700.0 * (10.0 ** (m / 2595.0) - 1.0)  # audio.py:36-38


def mel_filterbank(n_fft=N_FFT, n_mels=N_MELS, sr=SAMPLE_RATE, f_min=0.0, f_max=8000.0) -> np.ndarray:
    """audio.py:41-80, row by row like the reference."""
    n_freqs = n_fft // 2 + 1
    fft_freqs = np.linspace(0.0, sr / 2.0, n_freqs)
    pts = mel_to_hz(np.linspace(hz_to_mel(f_min), hz_to_mel(f_max), n_mels + 2))
    fb = np.zeros((n_mels, n_freqs), dtype=np.float32)
    for i in range(n_mels):
        lo, mid, hi = pts[i], pts[i + 1], pts[i + 2]
        fb[i] = np.maximum(0.0, np.minimum((fft_freqs - lo) / (mid - lo), (hi - fft_freqs) / (hi - mid)))
        if hi - lo > 0.0:
            fb[i] /= hi - lo
    return fb


_FB = None


def _fb():
    global _FB
    if _FB is None:
        _FB = mel_filterbank()
    return _FB


def _frames(audio: np.ndarray):
    window = np.hanning(N_FFT).astype(np.float32)  # audio.py:222
    padded = np.pad(audio, N_FFT // 2, mode="reflect")  # audio.py:223-224
    n_frames = 1 + (len(padded) - N_FFT) // HOP
    return window, padded, n_frames


def stft(audio: np.ndarray) -> np.ndarray:
    """audio.py:211-235 (per-frame loop)."""
    window, padded, n_frames = _frames(audio)
    out = np.empty((N_FFT // 2 + 1, n_frames), dtype=np.complex64)
    for i in range(n_frames):
        out[:, i] = np.fft.rfft(padded[i * HOP: i * HOP + N_FFT] * window, n=N_FFT)
    return out


def stft_fast(audio: np.ndarray) -> np.ndarray:
    window, padded, n_frames = _frames(audio)
    idx = np.arange(n_frames)[:, None] * HOP + np.arange(N_FFT)[None, :]
    return np.fft.rfft(padded[idx] * window[None, :], n=N_FFT, axis=1).T.astype(np.complex64)


def _finish(spec: np.ndarray) -> np.ndarray:
    power = np.abs(spec[:, :-1]) ** 2  # audio.py:266
    mel = _fb() @ power  # audio.py:272
    if mel.size == 0:
        raise ValueError("zero-size array to reduction operation maximum which has no identity")
    log_spec = np.log10(np.maximum(mel, 1e-10))  # audio.py:274
    log_spec = np.maximum(log_spec, log_spec.max() - 8.0)  # audio.py:275
    return ((log_spec + 4.0) / 4.0).astype(np.float32)  # audio.py:276


def log_mel_spectrogram(audio: np.ndarray) -> np.ndarray:
    return _finish(stft(np.asarray(audio, dtype=np.float32)))


def log_mel_spectrogram_fast(audio: np.ndarray) -> np.ndarray:
    return _finish(stft_fast(np.asarray(audio, dtype=np.float32)))
