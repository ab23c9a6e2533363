"""Audio-encoder configuration (mirrors AudioEncoderConfig, reference src/qwen3_asr_mlx/config.py:14-58)."""
from __future__ import annotations

import json
from dataclasses import dataclass
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
        """Read ``config.json`` from a local model directory and apply ``from_dict`` to it, as
        ModelConfig.from_pretrained does in the reference (config.py:130-150).  Hub download is
        out of scope here (no network); pass a local directory."""
        path = Path(model_path)
        if not path.is_dir():
            raise FileNotFoundError(f"{model_path} is not a local model directory (hub download is not supported)")
        d = json.loads((path / "config.json").read_text(encoding="utf-8"))
        return cls.from_dict(d)
