"""Audio-encoder and text-decoder configuration (mirror AudioEncoderConfig / TextDecoderConfig, reference
src/qwen3_asr_mlx/config.py:14-58, 61-100)."""
from __future__ import annotations

import json
from dataclasses import dataclass, field
from pathlib import Path
from typing import Any


@dataclass
class AudioEncoderConfig:
    """Hyper-parameters of the Qwen3-ASR audio encoder; defaults are the 1.7B model."""

    d_model: int = 1024
    encoder_layers: int = 24
    encoder_attention_heads: int = 16
    encoder_ffn_dim: int = 4096
    num_mel_bins: int = 128
    max_source_positions: int = 1500
    output_dim: int = 2048
    n_window: int = 50
    n_window_infer: int = 800
    conv_chunksize: int = 500
    activation_function: str = "gelu"
    downsample_hidden_size: int = 480

    @classmethod
    def from_dict(cls, d: dict[str, Any]) -> "AudioEncoderConfig":
        """Same lookup rules as the reference (config.py:31-58): nested ``audio_encoder_config``
        sub-dict if present, ``num_hidden_layers`` as fallback for ``encoder_layers``."""
        a = d.get("audio_encoder_config", d)
        defaults = cls()
        kw = {}
        for name in (
            "d_model", "encoder_attention_heads", "encoder_ffn_dim", "num_mel_bins", "max_source_positions",
            "output_dim", "n_window", "n_window_infer", "conv_chunksize", "activation_function",
            "downsample_hidden_size",
        ):
            kw[name] = a.get(name, getattr(defaults, name))
        kw["encoder_layers"] = a.get("encoder_layers", a.get("num_hidden_layers", defaults.encoder_layers))
        return cls(**kw)

    @classmethod
    def from_pretrained(cls, model_path: str | Path) -> "AudioEncoderConfig":
        """Read ``config.json`` from a local model directory, or from the hub when ``model_path`` is a repo id, and
        apply ``from_dict`` to it, as ModelConfig.from_pretrained does in the reference (config.py:130-150)."""
        from ._hub import config_dict

        return cls.from_dict(config_dict(model_path))


@dataclass
class TextDecoderConfig:
    """Hyper-parameters of the Qwen3 text decoder (reference config.py:61-76); defaults are the 1.7B model.
    ``mrope_section`` / ``rope_interleaved`` are carried but unused, as in the reference decoder (plain RoPE)."""

    hidden_size: int = 2048
    num_hidden_layers: int = 28
    num_attention_heads: int = 16
    num_key_value_heads: int = 8
    head_dim: int = 128
    intermediate_size: int = 6144
    hidden_act: str = "silu"
    vocab_size: int = 151936
    max_position_embeddings: int = 65536
    rms_norm_eps: float = 1e-6
    rope_theta: float = 1_000_000.0
    mrope_section: list = field(default_factory=lambda: [24, 20, 20])
    rope_interleaved: bool = True

    @classmethod
    def from_dict(cls, d: dict[str, Any]) -> "TextDecoderConfig":
        """Top-level keys of config.json, defaults for anything absent (reference config.py:78-100)."""
        defaults = cls()
        names = ("hidden_size", "num_hidden_layers", "num_attention_heads", "num_key_value_heads", "head_dim", "intermediate_size",
                 "hidden_act", "vocab_size", "max_position_embeddings", "rms_norm_eps", "rope_theta", "mrope_section", "rope_interleaved")
        return cls(**{n: d.get(n, getattr(defaults, n)) for n in names})

    @classmethod
    def from_pretrained(cls, model_path: str | Path) -> "TextDecoderConfig":
        from ._hub import config_dict

        return cls.from_dict(config_dict(model_path))


@dataclass
class ModelConfig:
    """Top-level Qwen3-ASR configuration (reference config.py:103-150): both sub-configs plus the audio special-token ids."""

    audio_encoder: AudioEncoderConfig = field(default_factory=AudioEncoderConfig)
    text_decoder: TextDecoderConfig = field(default_factory=TextDecoderConfig)
    audio_token_id: int = 151676
    audio_start_token_id: int = 151669
    audio_end_token_id: int = 151670

    @classmethod
    def from_dict(cls, d: dict[str, Any]) -> "ModelConfig":
        return cls(audio_encoder=AudioEncoderConfig.from_dict(d), text_decoder=TextDecoderConfig.from_dict(d),
                   audio_token_id=d.get("audio_token_id", 151676), audio_start_token_id=d.get("audio_start_token_id", 151669),
                   audio_end_token_id=d.get("audio_end_token_id", 151670))

    @classmethod
    def from_pretrained(cls, model_path: str | Path) -> "ModelConfig":
        """``config.json`` of a local model directory or of a hub repo id (reference config.py:130-150)."""
        from ._hub import config_dict

        return cls.from_dict(config_dict(model_path))
