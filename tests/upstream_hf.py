"""Third, independent encoder oracle (TEST INFRASTRUCTURE): the model authors' own PyTorch
implementation of this audio tower, `Qwen3OmniMoeAudioEncoder` from the `transformers` package
installed in this image (transformers 5.5.0, models/qwen3_omni_moe/modeling_qwen3_omni_moe.py:625-835).

Qwen3-ASR re-uses the Qwen3-Omni "AuT" audio tower; the reference's encoder.py is an MLX port of it
and keeps the same parameter names, so a strict `load_state_dict` of our `audio_tower.*` tensors
(Conv2d weights moved from the MLX layout (O,kH,kW,I) to torch's (O,I,kH,kW)) succeeds.  Agreement of
oracle/encoder_{np,torch}.py with this implementation pins the [MLX-semantics] assumptions of
SURVEY.md §8c (NHWC cross-correlation, flatten order channel*16+freq, exact-erf GELU, LayerNorm
eps 1e-5, PE restart per chunk, 104-token windows) against code that neither this repo nor the
reference wrote.

Two known, deliberate differences between upstream and the reference, both avoided by the tests:
  * upstream pads chunks to the longest chunk *in the batch* (`pad_sequence`), the reference always
    to 100 frames (encoder.py:262-266) — identical whenever the utterance has >= 1 full chunk;
  * upstream's eager attention ignores `cu_seqlens` (only its FlashAttention-2 varlen path honours
    them, modeling_qwen3_omni_moe.py:676-683); `register_varlen_attention` supplies a plain-torch
    attention that does what the FA2 path does.
"""
from __future__ import annotations

import numpy as np
import torch


def register_varlen_attention(name: str = "qasr_varlen_eager") -> str:
    from transformers import AttentionInterface

    def varlen_eager(module, query, key, value, attention_mask=None, scaling=None, cu_seq_lens_q=None, **kw):
        # query/key/value: (1, H, n, Dh); block-diagonal attention over cu_seq_lens segments (FA2 varlen semantics)
        out = torch.empty_like(query)
        cu = [int(v) for v in cu_seq_lens_q]
        for s, e in zip(cu[:-1], cu[1:]):
            w = torch.softmax((query[:, :, s:e] @ key[:, :, s:e].transpose(-1, -2)) * scaling, dim=-1)
            out[:, :, s:e] = w @ value[:, :, s:e]
        return out.transpose(1, 2).contiguous(), None  # (1, n, H, Dh) like the library interfaces

    AttentionInterface.register(name, varlen_eager)
    return name


def build_upstream(cfg, params):
    """Instantiate the upstream module for our AudioEncoderConfig and load our parameter dict (strict)."""
    from transformers.models.qwen3_omni_moe.configuration_qwen3_omni_moe import Qwen3OmniMoeAudioEncoderConfig
    from transformers.models.qwen3_omni_moe.modeling_qwen3_omni_moe import Qwen3OmniMoeAudioEncoder

    hc = Qwen3OmniMoeAudioEncoderConfig(
        num_mel_bins=cfg.num_mel_bins, encoder_layers=cfg.encoder_layers, encoder_attention_heads=cfg.encoder_attention_heads,
        encoder_ffn_dim=cfg.encoder_ffn_dim, d_model=cfg.d_model, max_source_positions=cfg.max_source_positions,
        n_window=cfg.n_window, n_window_infer=cfg.n_window_infer, output_dim=cfg.output_dim,
        downsample_hidden_size=cfg.downsample_hidden_size, activation_function="gelu", conv_chunksize=500)
    hc._attn_implementation = register_varlen_attention()
    with torch.device("meta"):
        model = Qwen3OmniMoeAudioEncoder(hc)
    model = model.to_empty(device="cpu").eval()
    sd = {}
    for k, v in params.items():
        t = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
        if k.startswith("conv2d") and k.endswith(".weight"):
            t = t.permute(0, 3, 1, 2).contiguous()  # MLX (O,kH,kW,I) -> torch (O,I,kH,kW)
        sd[k] = t
    result = model.load_state_dict(sd, strict=True)  # raises if any name or shape differs
    assert not result.missing_keys and not result.unexpected_keys
    # non-persistent buffer: rebuild after to_empty
    from transformers.models.qwen3_omni_moe.modeling_qwen3_omni_moe import SinusoidsPositionEmbedding
    model.positional_embedding = SinusoidsPositionEmbedding(cfg.max_source_positions, cfg.d_model)
    return model


@torch.no_grad()
def upstream_forward(model, mel) -> np.ndarray:
    mel = torch.as_tensor(np.asarray(mel, dtype=np.float32))
    return model(mel, feature_lens=torch.tensor([mel.shape[1]])).last_hidden_state.numpy()


def build_upstream_decoder(cfg, params):
    """The model authors' Qwen3 decoder (transformers models/qwen3/modeling_qwen3.py: GQA, per-head q/k RMSNorm, rotate-half
    RoPE, SwiGLU, tied lm_head) loaded with our parameter dict; key names match the checkpoint's ``model.*`` names."""
    from transformers import Qwen3Config, Qwen3ForCausalLM

    hc = Qwen3Config(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                     num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                     num_key_value_heads=cfg.num_key_value_heads, head_dim=cfg.head_dim, hidden_act="silu",
                     max_position_embeddings=cfg.max_position_embeddings, rms_norm_eps=cfg.rms_norm_eps, rope_theta=cfg.rope_theta,
                     tie_word_embeddings=True, attention_bias=False, use_sliding_window=False)
    hc._attn_implementation = "eager"
    model = Qwen3ForCausalLM(hc).eval().float()
    sd = {"model." + k: (v.detach().float().cpu() if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v, dtype=np.float32)))
          for k, v in params.items()}
    sd["lm_head.weight"] = sd["model.embed_tokens.weight"]
    result = model.load_state_dict(sd, strict=True)
    assert not result.missing_keys and not result.unexpected_keys
    return model


@torch.no_grad()
def upstream_decoder_forward(model, embeddings):
    emb = torch.as_tensor(np.asarray(embeddings, dtype=np.float32)) if not isinstance(embeddings, torch.Tensor) else embeddings.float().cpu()
    out = model(inputs_embeds=emb[None], use_cache=True)
    keys = torch.stack([layer_kv[0][0] for layer_kv in _cache_layers(out.past_key_values)]).numpy()
    values = torch.stack([layer_kv[1][0] for layer_kv in _cache_layers(out.past_key_values)]).numpy()
    return out.logits[0].numpy(), keys, values


def _cache_layers(cache):
    if hasattr(cache, "layers"):
        return [(l.keys, l.values) for l in cache.layers]
    return list(cache)
