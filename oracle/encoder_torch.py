"""torch-CPU fp32 restatement of the reference audio encoder (TEST INFRASTRUCTURE).

Follows /root/reference/src/qwen3_asr_mlx/encoder.py with library ops:
  positional table :21-44 | attention :51-86 | layer :93-122 | module shapes :142-191 |
  length + window rules :197-229 | forward, in this order of operations :235-323.
MLX semantics encoded here: Linear y = x W^T + b with W (out,in); Conv2d is NHWC
cross-correlation with weights (O,kH,kW,I) and zero padding; LayerNorm eps 1e-5, biased
variance, affine; nn.gelu is the exact erf form; SDPA = softmax(scale q k^T + mask) v.
Attention is evaluated per window (equivalent to the reference's -1e9 block mask in fp32:
exp(-1e9 - max) underflows to exactly 0).
PINNED: tests/test_reference_pin.py compares this restatement with tests/golden/encoder_reference.npz, the outputs of
the reference's own encoder.py executed unmodified (oracle/reference_ref.py): <= 1e-5 relative, measured 5e-7.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F


def conv_output_length(n: int) -> int:  # encoder.py:197-207
    for _ in range(3):
        n = (n - 1) // 2 + 1
    return n


def positional_table(rows: int, d_model: int) -> torch.Tensor:  # encoder.py:29-40
    half = d_model // 2
    log_timescale = math.log(10000.0) / (half - 1)
    inv = torch.exp(-torch.arange(half, dtype=torch.float32) * log_timescale)
    scaled = torch.arange(rows, dtype=torch.float32)[:, None] * inv[None, :]
    return torch.cat([torch.sin(scaled), torch.cos(scaled)], dim=1)


def _t(params: Dict[str, np.ndarray], name: str) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(params[name], dtype=np.float32))


def _linear(x, params, prefix, bias=True):
    return F.linear(x, _t(params, prefix + ".weight"), _t(params, prefix + ".bias") if bias else None)


def _ln(x, params, prefix):
    return F.layer_norm(x, (x.shape[-1],), _t(params, prefix + ".weight"), _t(params, prefix + ".bias"), eps=1e-5)


def _conv(x_nchw, params, prefix):
    w = _t(params, prefix + ".weight").permute(0, 3, 1, 2).contiguous()  # (O,kH,kW,I) -> (O,I,kH,kW)
    return F.gelu(F.conv2d(x_nchw, w, _t(params, prefix + ".bias"), stride=2, padding=1))


@torch.no_grad()
def encoder_forward(params: Dict[str, np.ndarray], cfg, mel, return_intermediates: bool = False, n_layers: Optional[int] = None):
    """mel (128, T) float32 -> (n_tokens, output_dim) float32 (the reference adds a leading 1)."""
    mel = torch.as_tensor(np.asarray(mel, dtype=np.float32))
    if mel.ndim == 3:
        mel = mel[0]  # encoder.py:249-250
    n_mels, T = mel.shape
    chunk = cfg.n_window * 2
    # chunking with zero padding of the last chunk (encoder.py:258-268)
    n_chunks = (T + chunk - 1) // chunk
    padded = torch.zeros((n_mels, n_chunks * chunk), dtype=torch.float32)
    padded[:, :T] = mel
    real = [min(chunk, T - i * chunk) for i in range(n_chunks)]
    x = padded.reshape(n_mels, n_chunks, chunk).permute(1, 0, 2)[:, None]  # (c,1,128,100) NCHW, H=mel W=time
    x = _conv(x, params, "conv2d1")
    x = _conv(x, params, "conv2d2")
    x = _conv(x, params, "conv2d3")  # (c, C, freq, time)
    c, C, Fq, Tt = x.shape
    x = x.permute(0, 3, 1, 2).reshape(c, Tt, C * Fq)  # flat index = channel*freq_bins + freq (encoder.py:277-278)
    x = _linear(x, params, "conv_out", bias=False)
    x = x + positional_table(Tt, cfg.d_model)[None]  # encoder.py:284-286 (before stripping)
    hidden = torch.cat([x[i, : conv_output_length(real[i])] for i in range(c)], dim=0)  # encoder.py:289-293
    inter = {"stem": hidden.clone()}
    n = hidden.shape[0]
    window = Tt * (cfg.n_window_infer // chunk)  # encoder.py:298-300
    H = cfg.encoder_attention_heads
    Dh = cfg.d_model // H
    layers = cfg.encoder_layers if n_layers is None else n_layers
    for li in range(layers):
        p = f"layers.{li}."
        y = _ln(hidden, params, p + "self_attn_layer_norm")
        q = _linear(y, params, p + "self_attn.q_proj").reshape(n, H, Dh).transpose(0, 1)
        k = _linear(y, params, p + "self_attn.k_proj").reshape(n, H, Dh).transpose(0, 1)
        v = _linear(y, params, p + "self_attn.v_proj").reshape(n, H, Dh).transpose(0, 1)
        att = torch.empty_like(q)
        for s in range(0, n, window):
            e = min(n, s + window)
            w = torch.softmax((q[:, s:e] @ k[:, s:e].transpose(1, 2)) * (Dh ** -0.5), dim=-1)
            att[:, s:e] = w @ v[:, s:e]
        hidden = hidden + _linear(att.transpose(0, 1).reshape(n, H * Dh), params, p + "self_attn.out_proj")
        y = _ln(hidden, params, p + "final_layer_norm")
        hidden = hidden + _linear(F.gelu(_linear(y, params, p + "fc1")), params, p + "fc2")
        if li == 0:
            inter["layer0"] = hidden.clone()
    inter["hidden"] = hidden.clone()
    out = _linear(F.gelu(_linear(_ln(hidden, params, "ln_post"), params, "proj1")), params, "proj2")
    if return_intermediates:
        return out.numpy(), {k_: v_.numpy() for k_, v_ in inter.items()}
    return out.numpy()
