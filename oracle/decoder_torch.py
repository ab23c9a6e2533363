"""torch-CPU fp32 restatement of the reference text decoder's PREFILL forward (TEST INFRASTRUCTURE).

Follows /root/reference/src/qwen3_asr_mlx/decoder.py with library ops:
  causal mask :68-81 | SwiGLU MLP :88-99 | attention (q/k/v_proj, q_norm/k_norm, RoPE, cache, SDPA, o_proj) :106-178 |
  pre-norm layer :181-200 | TextDecoder.__call__ (layers, final norm, tied lm_head) :223-253,
and the prefill call site generate.py:266-275 (cache empty, offset 0).
MLX semantics encoded here: Linear y = x W^T (no biases in this module); nn.RMSNorm(dims, eps) = x * rsqrt(mean(x^2) + eps) * w;
nn.RoPE(dims, traditional=False, base): frequencies base^(-2i/dims), pairs (i, i + dims/2) ("rotate half"), position = offset + t;
mx.fast.scaled_dot_product_attention broadcasts each KV head over n_heads / n_kv_heads query heads (GQA) and applies
softmax(scale q k^T + mask) v with an fp32 softmax; the additive -1e9 causal mask equals masking in fp32.
PINNED: tests/test_reference_pin.py compares logits / cached keys / values with tests/golden/decoder_reference.npz, the
outputs of the reference's own decoder.py executed unmodified (oracle/reference_ref.py), <= 1e-5; tests/test_oracle_decoder.py
additionally pins it against transformers' Qwen3 implementation (the model authors' code).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F


def _w(params, name) -> torch.Tensor:
    v = params[name]
    return v.detach().float().cpu() if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))


def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * w


def rope(x: torch.Tensor, theta: float, offset: int = 0) -> torch.Tensor:
    """x: (heads, T, head_dim); nn.RoPE(head_dim, traditional=False, base=theta) (decoder.py:128,164-166)."""
    _, T, D = x.shape
    half = D // 2
    inv = theta ** (-torch.arange(half, dtype=torch.float32) / half)
    ang = (torch.arange(T, dtype=torch.float32) + offset)[:, None] * inv[None, :]
    cos, sin = torch.cos(ang), torch.sin(ang)
    x1, x2 = x[..., :half], x[..., half:]
    return torch.cat([x1 * cos - x2 * sin, x1 * sin + x2 * cos], dim=-1)


@torch.no_grad()
def decoder_prefill(params: Dict[str, object], cfg, embeddings, n_layers: Optional[int] = None):
    """One prompt: embeddings (T, hidden) -> dict(logits (T, vocab), hidden (T, hidden), keys / values (L, n_kv, T, head_dim))."""
    h = torch.as_tensor(np.asarray(embeddings, dtype=np.float32)) if not isinstance(embeddings, torch.Tensor) else embeddings.detach().float().cpu()
    T = h.shape[0]
    Hq, Hkv, D = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
    eps, scale = cfg.rms_norm_eps, cfg.head_dim ** -0.5
    mask = torch.triu(torch.full((T, T), -1e9), diagonal=1)  # decoder.py:68-81 with offset 0
    keys, values = [], []
    L = cfg.num_hidden_layers if n_layers is None else n_layers
    for i in range(L):
        p = f"layers.{i}."
        x = rms_norm(h, _w(params, p + "input_layernorm.weight"), eps)
        q = F.linear(x, _w(params, p + "self_attn.q_proj.weight")).reshape(T, Hq, D)
        k = F.linear(x, _w(params, p + "self_attn.k_proj.weight")).reshape(T, Hkv, D)
        v = F.linear(x, _w(params, p + "self_attn.v_proj.weight")).reshape(T, Hkv, D)
        q = rms_norm(q, _w(params, p + "self_attn.q_norm.weight"), eps).transpose(0, 1)  # (Hq, T, D)
        k = rms_norm(k, _w(params, p + "self_attn.k_norm.weight"), eps).transpose(0, 1)
        v = v.transpose(0, 1)
        q, k = rope(q, cfg.rope_theta), rope(k, cfg.rope_theta)
        keys.append(k.clone())
        values.append(v.clone())
        g = Hq // Hkv
        kk, vv = k.repeat_interleave(g, dim=0), v.repeat_interleave(g, dim=0)  # query head j uses kv head j // g
        att = torch.softmax(q @ kk.transpose(1, 2) * scale + mask, dim=-1) @ vv  # (Hq, T, D)
        h = h + F.linear(att.transpose(0, 1).reshape(T, Hq * D), _w(params, p + "self_attn.o_proj.weight"))
        x = rms_norm(h, _w(params, p + "post_attention_layernorm.weight"), eps)
        gate = F.linear(x, _w(params, p + "mlp.gate_proj.weight"))
        up = F.linear(x, _w(params, p + "mlp.up_proj.weight"))
        h = h + F.linear(F.silu(gate) * up, _w(params, p + "mlp.down_proj.weight"))
    out = {"hidden": h.numpy()}
    xn = rms_norm(h, _w(params, "norm.weight"), eps)
    out["logits"] = (xn @ _w(params, "embed_tokens.weight").T).numpy()  # tied lm_head, decoder.py:252
    out["keys"] = torch.stack(keys).numpy() if keys else np.zeros((0, Hkv, T, D), np.float32)
    out["values"] = torch.stack(values).numpy() if values else np.zeros((0, Hkv, T, D), np.float32)
    return out


def decoder_prefill_batch(params, cfg, embeddings, seq_offsets: Sequence[int]):
    """The reference runs one prompt at a time: a batch is a loop."""
    return [decoder_prefill(params, cfg, embeddings[int(a): int(b)]) for a, b in zip(seq_offsets[:-1], seq_offsets[1:])]
