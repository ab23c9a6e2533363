"""Encoder parameter inventory, seeded random initialisation and safetensors I/O.

Parameter names and shapes follow the reference's module attributes
(src/qwen3_asr_mlx/encoder.py:60-63,98-104,148-191), i.e. the keys of ``model.safetensors``
after the ``audio_tower.`` prefix is stripped (encoder.py:349-356):
Linear weights are (out, in); Conv2d weights are (O, kH, kW, I); LayerNorm has weight and bias.

Random initialisation reproduces the *distributions* MLX uses for freshly constructed modules
(Linear: W, b ~ U(+-1/sqrt(in)); Conv2d: W ~ U(+-1/sqrt(in*kh*kw)), b = 0; LayerNorm: w = 1, b = 0),
drawn from numpy's PCG64 in a fixed key order, so the same seed gives the same tensors for the
CUDA path and for the CPU oracle.  ``exercise_all=True`` additionally randomises the parameters
MLX initialises to constants (conv biases, LayerNorm affine), so parity tests cover them too.
"""
from __future__ import annotations

from pathlib import Path
from typing import Dict, Iterator, Tuple

import numpy as np

from .config import AudioEncoderConfig

PREFIX = "audio_tower."


def parameter_shapes(cfg: AudioEncoderConfig) -> Iterator[Tuple[str, Tuple[int, ...]]]:
    """Yield (name, shape) for every encoder parameter, in a fixed order."""
    C, D, F, O = cfg.downsample_hidden_size, cfg.d_model, cfg.encoder_ffn_dim, cfg.output_dim
    freq = ((((cfg.num_mel_bins + 1) // 2) + 1) // 2 + 1) // 2  # encoder.py:171-173
    yield "conv2d1.weight", (C, 3, 3, 1)
    yield "conv2d1.bias", (C,)
    yield "conv2d2.weight", (C, 3, 3, C)
    yield "conv2d2.bias", (C,)
    yield "conv2d3.weight", (C, 3, 3, C)
    yield "conv2d3.bias", (C,)
    yield "conv_out.weight", (D, C * freq)
    for i in range(cfg.encoder_layers):
        p = f"layers.{i}."
        yield p + "self_attn_layer_norm.weight", (D,)
        yield p + "self_attn_layer_norm.bias", (D,)
        for proj in ("q_proj", "k_proj", "v_proj", "out_proj"):
            yield p + f"self_attn.{proj}.weight", (D, D)
            yield p + f"self_attn.{proj}.bias", (D,)
        yield p + "final_layer_norm.weight", (D,)
        yield p + "final_layer_norm.bias", (D,)
        yield p + "fc1.weight", (F, D)
        yield p + "fc1.bias", (F,)
        yield p + "fc2.weight", (D, F)
        yield p + "fc2.bias", (D,)
    yield "ln_post.weight", (D,)
    yield "ln_post.bias", (D,)
    yield "proj1.weight", (D, D)
    yield "proj1.bias", (D,)
    yield "proj2.weight", (O, D)
    yield "proj2.bias", (O,)


def random_init(cfg: AudioEncoderConfig, seed: int = 1234, exercise_all: bool = False) -> Dict[str, np.ndarray]:
    """Seeded float32 parameters with MLX's default initialisation distributions."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out: Dict[str, np.ndarray] = {}
    linear_fan_in = 1
    for name, shape in parameter_shapes(cfg):
        leaf = name.rsplit(".", 1)[-1]
        is_ln = "layer_norm" in name or name.startswith("ln_post")
        if is_ln:
            if exercise_all:
                v = (1.0 + 0.2 * rng.standard_normal(shape)) if leaf == "weight" else 0.1 * rng.standard_normal(shape)
            else:
                v = np.ones(shape) if leaf == "weight" else np.zeros(shape)
        elif name.startswith("conv2d"):
            fan_in = shape[1] * shape[2] * shape[3] if leaf == "weight" else None
            if leaf == "weight":
                s = 1.0 / np.sqrt(fan_in)
                v = rng.uniform(-s, s, size=shape)
            else:
                v = rng.uniform(-0.05, 0.05, size=shape) if exercise_all else np.zeros(shape)
        else:  # Linear
            if leaf == "weight":
                s = 1.0 / np.sqrt(shape[1])
                v = rng.uniform(-s, s, size=shape)
                linear_fan_in = shape[1]  # the bias of the same Linear follows and shares the bound
            else:
                s = 1.0 / np.sqrt(linear_fan_in)
                v = rng.uniform(-s, s, size=shape)
        out[name] = np.ascontiguousarray(v, dtype=np.float32)
    return out


def save_safetensors(params: Dict[str, np.ndarray], path: str | Path, prefix: str = PREFIX) -> None:
    """Write ``model.safetensors`` with the reference's key naming (``audio_tower.`` prefix)."""
    from safetensors.numpy import save_file

    save_file({prefix + k: np.ascontiguousarray(v) for k, v in params.items()}, str(path))


def load_safetensors(path: str | Path, prefix: str = PREFIX) -> Dict[str, np.ndarray]:
    """Read the ``audio_tower.*`` tensors of a safetensors file and strip the prefix
    (reference load_encoder_weights, encoder.py:347-356; no transposes: MLX layout)."""
    from safetensors import safe_open

    out: Dict[str, np.ndarray] = {}
    with safe_open(str(path), framework="pt") as f:  # "pt" so that bf16 checkpoints load
        for key in f.keys():
            if not key.startswith(prefix):
                continue
            t = f.get_tensor(key)
            out[key[len(prefix):]] = t.float().numpy() if str(t.dtype) == "torch.bfloat16" else t.numpy()
    return out
