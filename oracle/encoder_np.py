"""numpy float64 restatement of the reference audio encoder with explicit loops (TEST INFRASTRUCTURE).

Independent of oracle/encoder_torch.py on purpose: no library convolution (3x3 patches are
gathered by hand), hand-written LayerNorm / softmax, scipy's erf, and — like the reference
itself (encoder.py:209-229, 311) — a DENSE additive -1e9 block mask over all n tokens instead of
per-window attention.  Use on small inputs only (O(n^2) mask, python loops).
Reference lines are the same as listed in encoder_torch.py.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
from scipy.special import erf


def conv_output_length(n: int) -> int:
    for _ in range(3):
        n = (n - 1) // 2 + 1
    return n


def gelu(x):
    return 0.5 * x * (1.0 + erf(x / math.sqrt(2.0)))


def layer_norm(x, w, b, eps=1e-5):
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * w + b


def conv2d_nhwc_s2(x, w, b):
    """x (B,H,W,I), w (O,3,3,I) cross-correlation, stride 2, zero padding 1 -> (B,Ho,Wo,O)."""
    B, H, W, I = x.shape
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    xp = np.zeros((B, H + 2, W + 2, I), dtype=np.float64)
    xp[:, 1:H + 1, 1:W + 1] = x
    out = np.zeros((B, Ho, Wo, w.shape[0]), dtype=np.float64)
    for kh in range(3):
        for kw in range(3):
            patch = xp[:, kh:kh + 2 * Ho:2, kw:kw + 2 * Wo:2, :]  # (B,Ho,Wo,I)
            out += patch @ w[:, kh, kw, :].T
    return out + b


def positional_table(rows: int, d_model: int) -> np.ndarray:
    half = d_model // 2
    inv = np.exp(-np.arange(half, dtype=np.float64) * (math.log(10000.0) / (half - 1)))
    s = np.arange(rows, dtype=np.float64)[:, None] * inv[None, :]
    return np.concatenate([np.sin(s), np.cos(s)], axis=1)


def encoder_forward(params: Dict[str, np.ndarray], cfg, mel) -> np.ndarray:
    P = {k: np.asarray(v, dtype=np.float64) for k, v in params.items()}
    mel = np.asarray(mel, dtype=np.float64)
    if mel.ndim == 3:
        mel = mel[0]
    n_mels, T = mel.shape
    chunk = cfg.n_window * 2
    chunks, real = [], []
    for off in range(0, T, chunk):
        seg = mel[:, off:off + chunk]
        real.append(seg.shape[1])
        if seg.shape[1] < chunk:
            seg = np.concatenate([seg, np.zeros((n_mels, chunk - seg.shape[1]))], axis=1)
        chunks.append(seg)
    x = np.stack(chunks, axis=0)[:, :, :, None]  # (c,128,100,1)
    for name in ("conv2d1", "conv2d2", "conv2d3"):
        x = gelu(conv2d_nhwc_s2(x, P[name + ".weight"], P[name + ".bias"]))
    B, freq, time, ch = x.shape
    x = x.transpose(0, 2, 3, 1).reshape(B, time, ch * freq)
    x = x @ P["conv_out.weight"].T
    x = x + positional_table(time, cfg.d_model)[None]
    hidden = np.concatenate([x[i, :conv_output_length(real[i])] for i in range(B)], axis=0)
    n = hidden.shape[0]
    window = time * (cfg.n_window_infer // chunk)
    cu = [0]
    for _ in range(n // window):
        cu.append(cu[-1] + window)
    if n % window:
        cu.append(cu[-1] + n % window)
    mask = None
    if len(cu) > 2:
        mask = np.full((n, n), -1e9)
        for lo, hi in zip(cu[:-1], cu[1:]):
            mask[lo:hi, lo:hi] = 0.0
    H = cfg.encoder_attention_heads
    Dh = cfg.d_model // H
    for li in range(cfg.encoder_layers):
        p = f"layers.{li}."
        y = layer_norm(hidden, P[p + "self_attn_layer_norm.weight"], P[p + "self_attn_layer_norm.bias"])
        q = (y @ P[p + "self_attn.q_proj.weight"].T + P[p + "self_attn.q_proj.bias"]).reshape(n, H, Dh).transpose(1, 0, 2)
        k = (y @ P[p + "self_attn.k_proj.weight"].T + P[p + "self_attn.k_proj.bias"]).reshape(n, H, Dh).transpose(1, 0, 2)
        v = (y @ P[p + "self_attn.v_proj.weight"].T + P[p + "self_attn.v_proj.bias"]).reshape(n, H, Dh).transpose(1, 0, 2)
        s = q @ k.transpose(0, 2, 1) * (Dh ** -0.5)
        if mask is not None:
            s = s + mask[None]
        s = s - s.max(axis=-1, keepdims=True)
        w = np.exp(s)
        w /= w.sum(axis=-1, keepdims=True)
        att = (w @ v).transpose(1, 0, 2).reshape(n, H * Dh)
        hidden = hidden + att @ P[p + "self_attn.out_proj.weight"].T + P[p + "self_attn.out_proj.bias"]
        y = layer_norm(hidden, P[p + "final_layer_norm.weight"], P[p + "final_layer_norm.bias"])
        y = gelu(y @ P[p + "fc1.weight"].T + P[p + "fc1.bias"])
        hidden = hidden + y @ P[p + "fc2.weight"].T + P[p + "fc2.bias"]
    y = layer_norm(hidden, P["ln_post.weight"], P["ln_post.bias"])
    y = gelu(y @ P["proj1.weight"].T + P["proj1.bias"])
    return y @ P["proj2.weight"].T + P["proj2.bias"]
