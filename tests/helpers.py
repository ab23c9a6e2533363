"""Shared test utilities (signal recipe of SURVEY.md §8d, error metrics)."""
import numpy as np


def synth(rng, n):
    t = np.arange(n) / 16000.0
    x = 0.1 * rng.standard_normal(n)
    for _ in range(3):
        x += 0.3 * np.sin(2 * np.pi * rng.uniform(100, 4000) * t + rng.uniform(0, 6.28)) * (0.5 + 0.5 * np.sin(2 * np.pi * rng.uniform(0.1, 1.0) * t))
    return np.clip(x, -1, 1).astype(np.float32)


def rel_err(got, ref):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-30))


def bf16_round(a):
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) >> 16
    return (u.astype(np.uint32) << 16).view(np.float32)


def bf16_bits(a):
    return (bf16_round(a).view(np.uint32) >> 16).astype(np.uint16)


MEL_TOL = 1e-4   # BASELINE.json north_star: log-mel max-abs error <= 1e-4
EMB_TOL = 2e-2   # BASELINE.json north_star: encoder-embedding relative error <= 2e-2 (bf16 vs fp32)


def make_wav(seed, n, rate, channels, fmt):
    """Deterministic RIFF/WAVE bytes: fmt in {"pcm16", "pcm32", "f32"} (reference audio.py:103-170 fast path)."""
    import struct

    x = np.random.default_rng(seed).uniform(-0.9, 0.9, size=(n, channels))
    if fmt == "pcm16":
        body, code, bits = (x * 32767).astype("<i2").tobytes(), 1, 16
    elif fmt == "pcm32":
        body, code, bits = (x * 2147483647).astype("<i4").tobytes(), 1, 32
    else:
        body, code, bits = x.astype("<f4").tobytes(), 3, 32
    hdr = struct.pack("<HHIIHH", code, channels, rate, rate * channels * bits // 8, channels * bits // 8, bits)
    extra = b"LIST" + struct.pack("<I", 4) + b"abcd"  # an unrelated chunk the reader must skip
    chunks = b"fmt " + struct.pack("<I", len(hdr)) + hdr + extra + b"data" + struct.pack("<I", len(body)) + body
    return b"RIFF" + struct.pack("<I", 4 + len(chunks)) + b"WAVE" + chunks


WAV_CASES = {
    "pcm16_mono_16k": dict(seed=1, n=4000, rate=16000, channels=1, fmt="pcm16"),
    "pcm16_stereo_16k": dict(seed=2, n=3000, rate=16000, channels=2, fmt="pcm16"),
    "pcm32_mono_16k": dict(seed=3, n=2500, rate=16000, channels=1, fmt="pcm32"),
    "f32_mono_16k": dict(seed=4, n=2000, rate=16000, channels=1, fmt="f32"),
    "pcm16_mono_8k": dict(seed=5, n=2000, rate=8000, channels=1, fmt="pcm16"),        # upsampling by linear interpolation
    "f32_stereo_44k": dict(seed=6, n=4410, rate=44100, channels=2, fmt="f32"),        # downsampling
    "pcm16_mono_22k": dict(seed=7, n=2205, rate=22050, channels=1, fmt="pcm16"),
}
