"""The reference package itself, imported UNMODIFIED from /root/reference/src (authoring container only).

`import qwen3_asr_mlx` needs `mlx.core` / `mlx.nn`, which cannot be installed here; oracle/_mlx_shim provides a
torch-CPU fp32 stand-in for exactly the MLX surface the reference touches (see its headers for the semantics it
restates).  On top of it the reference's own `AudioEncoder.__call__` (encoder.py:235-323), `TextDecoder.__call__`
(decoder.py:223-253), `prepare_inputs` (generate.py:20-81), `build_prompt` (tokenizer.py:56-86) and `LANGUAGE_MAP`
(model.py:28-96) run verbatim.  oracle/gen_golden.py stores their outputs in tests/golden/*_reference.npz; those
fixtures pin oracle/encoder_torch.py, encoder_np.py, decoder_torch.py, prompt_np.py and the CUDA path.

/root/reference does not exist on the GPU box: nothing at run time there imports this module (tests that do are
skipped when `available()` is false).  TEST INFRASTRUCTURE: never imported by qwen3_asr_mlx_b200/.
"""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np

REFERENCE_SRC = "/root/reference/src"
_pkg = None


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "qwen3_asr_mlx"))


def package():
    """The imported reference package (`qwen3_asr_mlx`)."""
    global _pkg
    if _pkg is None:
        shim = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_mlx_shim")
        for p in (REFERENCE_SRC, shim):
            if p not in sys.path:
                sys.path.insert(0, p)
        _pkg = importlib.import_module("qwen3_asr_mlx")
    return _pkg


def _mx():
    package()
    return importlib.import_module("mlx.core")


def build_encoder(params, cfg):
    """Reference `AudioEncoder(config)` with `params` (name -> numpy fp32, names as in the checkpoint minus
    `audio_tower.`) loaded through the reference module's own strict `load_weights` (encoder.py:358)."""
    pkg, mx = package(), _mx()
    rcfg = pkg.AudioEncoderConfig(**{f: getattr(cfg, f) for f in pkg.AudioEncoderConfig.__dataclass_fields__ if hasattr(cfg, f)})
    enc = pkg.AudioEncoder(rcfg)
    enc.load_weights([(k, mx.array(np.asarray(v, dtype=np.float32))) for k, v in params.items()])
    return enc


def encoder_forward(params, cfg, mel, encoder=None):
    """mel (128, T) float32 -> (n_tokens, output_dim) float32: the reference's own forward, leading 1 dropped."""
    mx = _mx()
    enc = encoder if encoder is not None else build_encoder(params, cfg)
    out = enc(mx.array(np.asarray(mel, dtype=np.float32)))
    return np.asarray(out, dtype=np.float32)[0]


def build_decoder(params, cfg):
    pkg, mx = package(), _mx()
    rcfg = pkg.TextDecoderConfig(**{f: getattr(cfg, f) for f in pkg.TextDecoderConfig.__dataclass_fields__ if hasattr(cfg, f)})
    dec = pkg.TextDecoder(rcfg)
    dec.load_weights([(k, mx.array(np.asarray(v, dtype=np.float32))) for k, v in params.items()])
    return dec


def decoder_prefill(params, cfg, embeddings, decoder=None):
    """The prefill call of generate() (generate.py:266-275): `decoder(embeds, cache=KVCache(), is_embeds=True)`.
    Returns logits (T, vocab), keys / values (L, n_kv, T, head_dim) as held by the reference's KVCache."""
    pkg, mx = package(), _mx()
    dec = decoder if decoder is not None else build_decoder(params, cfg)
    cache = pkg.KVCache()
    logits = dec(mx.array(np.asarray(embeddings, dtype=np.float32)[None]), cache=cache, is_embeds=True)
    return {"logits": np.asarray(logits, dtype=np.float32)[0],
            "keys": np.stack([np.asarray(k, dtype=np.float32)[0] for k in cache.keys]),
            "values": np.stack([np.asarray(v, dtype=np.float32)[0] for v in cache.values])}


def prepare_inputs(encoder_output, input_ids, table):
    """generate.py:20-81 with an `nn.Embedding` whose weight is `table`."""
    mx = _mx()
    nn = importlib.import_module("mlx.nn")
    emb = nn.Embedding(table.shape[0], table.shape[1])
    emb.weight = mx.array(np.asarray(table, dtype=np.float32))
    out = package().prepare_inputs(mx.array(np.asarray(encoder_output, dtype=np.float32)), list(input_ids), emb)
    return np.asarray(out, dtype=np.float32)
