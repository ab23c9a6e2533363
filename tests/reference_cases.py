"""Seeded inputs shared by oracle/gen_golden.py (which runs the REFERENCE on them, authoring container only) and by the
CPU / GPU tests that compare the oracle restatements and the CUDA path with the stored reference outputs
(tests/golden/{encoder,decoder,prompt}_reference.npz).  Everything here is regenerated from seeds with numpy's PCG64 /
torch's CPU generator; each fixture also stores a float64 checksum of its input so a drifting generator is detected
instead of silently comparing different inputs."""
from __future__ import annotations

from typing import Dict

import numpy as np

from helpers import synth


def checksum(a) -> float:
    a = np.asarray(a, dtype=np.float64).ravel()
    return float(np.dot(a, np.cos(np.arange(a.size) * 0.37)))


# ----------------------------------------------------------------------------- encoder (reference encoder.py:235-323)
ENC_SMALL_FRAMES = (1, 50, 100, 250, 801, 927, 3000)   # 1 token .. 30 s; 801 -> 104 + 1 tokens, 927 -> tail chunk of 27 frames
ENC_WIDE_FRAMES = (100, 927)                           # 1.7B widths, 2 layers
CONFIG1_SAMPLES = 160000                               # BASELINE config 1: one 10 s utterance, full 1.7B architecture


def encoder_groups(which=("small", "wide2", "full")):
    from qwen3_asr_mlx_b200 import weights
    from qwen3_asr_mlx_b200.config import AudioEncoderConfig

    out = {}
    if "small" in which:
        cfg = AudioEncoderConfig(d_model=256, encoder_layers=2, encoder_attention_heads=4, encoder_ffn_dim=512, output_dim=256)
        out["small"] = (cfg, weights.random_init(cfg, seed=7, exercise_all=True))
    if "wide2" in which:
        cfg = AudioEncoderConfig(encoder_layers=2)
        out["wide2"] = (cfg, weights.random_init(cfg, seed=1234, exercise_all=True))
    if "full" in which:
        cfg = AudioEncoderConfig()
        out["full"] = (cfg, weights.random_init(cfg, seed=1234))
    return out


def _mel_of(seed: int, frames: int) -> np.ndarray:
    """A real log-mel block: the pinned mel oracle applied to the SURVEY §8d signal recipe."""
    from oracle import mel_np

    return mel_np.log_mel_spectrogram_fast(synth(np.random.default_rng(seed), 160 * frames + 37))


def encoder_inputs(group: str) -> Dict[str, np.ndarray]:
    if group == "small":
        return {f"T{t}": _mel_of(1000 + t, t) for t in ENC_SMALL_FRAMES}
    if group == "wide2":
        return {f"T{t}": _mel_of(2000 + t, t) for t in ENC_WIDE_FRAMES}
    if group == "full":
        from oracle import mel_np

        return {"config1_10s": mel_np.log_mel_spectrogram_fast(synth(np.random.default_rng(0), CONFIG1_SAMPLES))}
    raise KeyError(group)


# ----------------------------------------------------------------------------- decoder prefill (decoder.py:106-253)
DEC_SMALL_T = (1, 7, 65, 200)


def decoder_groups(which=("small", "wide1")):
    from qwen3_asr_mlx_b200 import decoder as dec
    from qwen3_asr_mlx_b200.config import TextDecoderConfig

    out = {}
    if "small" in which:
        cfg = TextDecoderConfig(hidden_size=256, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, intermediate_size=512, vocab_size=1024)
        out["small"] = (cfg, {k: v.numpy() for k, v in dec.random_init(cfg, seed=5, exercise_all=True).items()})
    if "wide1" in which:
        cfg = TextDecoderConfig(num_hidden_layers=1, vocab_size=2048)
        out["wide1"] = (cfg, {k: v.numpy() for k, v in dec.random_init(cfg, seed=9, exercise_all=True).items()})
    return out


def decoder_inputs(group: str) -> Dict[str, np.ndarray]:
    if group == "small":
        return {f"T{t}": np.random.default_rng(300 + t).standard_normal((t, 256)).astype(np.float32) for t in DEC_SMALL_T}
    if group == "wide1":
        return {"T90": np.random.default_rng(390).standard_normal((90, 2048)).astype(np.float32)}
    raise KeyError(group)


# ----------------------------------------------------------------------------- prompt assembly (generate.py:20-81)
def prompt_cases():
    """name -> (encoder_output (1, n, H) f32, language-name token ids or None, embedding table (V, H) f32)."""
    out = {}
    for name, (n_audio, lang, seed) in {"a37_lang2": (37, [6364, 100], 0), "a1_nolang": (1, None, 1), "a130_lang1": (130, [6364], 2)}.items():
        rng = np.random.default_rng(seed)
        table = rng.standard_normal((152000, 64)).astype(np.float32)
        audio = rng.standard_normal((1, n_audio, 64)).astype(np.float32)
        out[name] = (audio, lang, table)
    return out
