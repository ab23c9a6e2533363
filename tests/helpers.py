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
