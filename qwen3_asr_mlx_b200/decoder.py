"""Text-decoder PREFILL on the B200 (mirrors the reference's decoder module, src/qwen3_asr_mlx/decoder.py).

``TextDecoder`` / ``KVCache`` / ``load_decoder_weights`` keep the reference names.  What runs here is the prefill
forward of ``generate()`` (generate.py:266-275): the whole prompt (text embeddings with the audio embeddings of the
encoder scattered in) through the 28 decoder layers, filling the KV cache and producing the logits the first token is
sampled from.  It is batched: any number of prompts, varlen-packed, in one call; results equal a per-prompt loop.
The token-by-token loop (generate.py:289-313) is outside this path.  Compute runs in libqasr
(include/qasr_decoder.h); without the library or an sm_100 GPU every call raises (no CPU fallback).
"""
from __future__ import annotations

import ctypes
from pathlib import Path
from typing import Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, runtime
from ._array import DeviceArray
from .config import TextDecoderConfig

PREFIX = "model."


def parameter_shapes(cfg: TextDecoderConfig) -> Iterator[Tuple[str, Tuple[int, ...]]]:
    """(name, shape) of every decoder parameter, names as in the checkpoint minus ``model.`` (decoder.py:88-221)."""
    H, Q, KV, I = cfg.hidden_size, cfg.num_attention_heads * cfg.head_dim, cfg.num_key_value_heads * cfg.head_dim, cfg.intermediate_size
    yield "embed_tokens.weight", (cfg.vocab_size, H)
    for i in range(cfg.num_hidden_layers):
        p = f"layers.{i}."
        yield p + "input_layernorm.weight", (H,)
        yield p + "self_attn.q_proj.weight", (Q, H)
        yield p + "self_attn.k_proj.weight", (KV, H)
        yield p + "self_attn.v_proj.weight", (KV, H)
        yield p + "self_attn.o_proj.weight", (H, Q)
        yield p + "self_attn.q_norm.weight", (cfg.head_dim,)
        yield p + "self_attn.k_norm.weight", (cfg.head_dim,)
        yield p + "post_attention_layernorm.weight", (H,)
        yield p + "mlp.gate_proj.weight", (I, H)
        yield p + "mlp.up_proj.weight", (I, H)
        yield p + "mlp.down_proj.weight", (H, I)
    yield "norm.weight", (H,)


def random_init(cfg: TextDecoderConfig, seed: int = 4321, device=None, exercise_all: bool = False) -> Dict[str, torch.Tensor]:
    """Seeded float32 parameters with MLX's default distributions (Linear: U(+-1/sqrt(in)); Embedding: N(0, 1/sqrt(dims));
    RMSNorm: ones), generated with torch's Philox stream on ``device`` (CPU by default) so that the CUDA path and the
    oracle can share them.  ``exercise_all`` also randomises the norm weights."""
    dev = torch.device("cpu") if device is None else torch.device(device)
    gen = torch.Generator(device=dev).manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    for name, shape in parameter_shapes(cfg):
        if name.endswith("norm.weight") or name.endswith("layernorm.weight"):
            w = torch.ones(shape, device=dev)
            if exercise_all:
                w = w + 0.2 * torch.randn(shape, generator=gen, device=dev)
        elif name == "embed_tokens.weight":
            w = torch.randn(shape, generator=gen, device=dev) * (shape[1] ** -0.5)
        else:
            s = shape[1] ** -0.5
            w = (torch.rand(shape, generator=gen, device=dev) * 2.0 - 1.0) * s
        out[name] = w.float()
    return out


class KVCache:
    """Keys / values written by the prefill (reference KVCache, decoder.py:20-61).

    ``keys`` / ``values``: bf16 CUDA tensors ``(num_layers, n_tokens, n_kv_heads, head_dim)``, token-major and
    varlen-packed; prompt ``u`` owns rows ``[seq_offsets[u], seq_offsets[u+1])``.  ``layer(i, u)`` returns that prompt's
    ``(1, n_kv_heads, T, head_dim)`` views, the shapes the reference holds.  ``offset`` is the prompt length per prompt."""

    def __init__(self, keys: torch.Tensor, values: torch.Tensor, seq_offsets: np.ndarray):
        self.keys, self.values, self.seq_offsets = keys, values, np.asarray(seq_offsets, dtype=np.int64)

    @property
    def offset(self) -> List[int]:
        return [int(v) for v in np.diff(self.seq_offsets)]

    def layer(self, layer_idx: int, prompt: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
        a, b = int(self.seq_offsets[prompt]), int(self.seq_offsets[prompt + 1])
        k = self.keys[layer_idx, a:b].permute(1, 0, 2).unsqueeze(0)
        v = self.values[layer_idx, a:b].permute(1, 0, 2).unsqueeze(0)
        return k, v


class TextDecoder:
    """``TextDecoder(config)`` with the reference's constructor; ``prefill`` is the batched hot path and
    ``__call__(embeddings, cache=None, is_embeds=True)`` the reference-shaped single-prompt forward (all logits)."""

    def __init__(self, config: Optional[TextDecoderConfig] = None, device: Optional[int] = None):
        runtime.require_cuda()
        self.config = config or TextDecoderConfig()
        self.lib = _lib.load()
        self.device = runtime.local_device() if device is None else int(device)
        self.torch_device = torch.device("cuda", self.device)
        c = self.config
        cc = _lib.QasrDecoderConfig(c.hidden_size, c.num_hidden_layers, c.num_attention_heads, c.num_key_value_heads, c.head_dim,
                                    c.intermediate_size, c.vocab_size, c.rms_norm_eps, c.rope_theta)
        self._d = ctypes.c_void_p()
        _lib.check(self.lib.qasr_decoder_create(self.device, ctypes.byref(cc), ctypes.byref(self._d)), None, decoder=True)
        self._ready = False

    # ------------------------------------------------------------------ lifetime
    def _check(self, rc: int) -> None:
        _lib.check(rc, self._d, decoder=True)

    @property
    def ptr(self) -> ctypes.c_void_p:
        if not self._d:
            raise _lib.QasrError("decoder already destroyed")
        return self._d

    def close(self) -> None:
        if getattr(self, "_d", None):
            self.lib.qasr_decoder_destroy(self._d)
            self._d = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def stats(self) -> Dict[str, int]:
        s = _lib.QasrStats()
        self._check(self.lib.qasr_decoder_get_stats(self.ptr, ctypes.byref(s)))
        return {"kernel_launches": s.kernel_launches, "workspace_bytes": s.workspace_bytes, "weight_bytes": s.weight_bytes}

    # ------------------------------------------------------------------ weights
    def load_weights(self, items: Iterable[Tuple[str, object]] | Dict[str, object]) -> None:
        """Install parameters by reference name (numpy arrays, CPU or CUDA torch tensors; float32 or bfloat16)."""
        if self._ready:
            raise _lib.QasrError("weights already loaded for this decoder; construct a new TextDecoder")
        pairs = items.items() if isinstance(items, dict) else items
        dts = {torch.float32: _lib.QASR_F32, torch.bfloat16: _lib.QASR_BF16}
        for name, value in pairs:
            if isinstance(value, DeviceArray):
                value = value.tensor
            if isinstance(value, torch.Tensor):
                t = value.detach()
                if t.dtype not in dts:
                    t = t.float()
                t = t.contiguous()
                if t.is_cuda and t.device != self.torch_device:
                    t = t.to(self.torch_device)
                dtype = dts[t.dtype] | (_lib.QASR_DEVICE_PTR if t.is_cuda else 0)
                ptr, shape = t.data_ptr(), tuple(t.shape)
                if t.is_cuda:
                    torch.cuda.synchronize(self.torch_device)  # libqasr reads it on the default stream
            else:
                arr = np.ascontiguousarray(np.asarray(value), dtype=np.float32)
                t, dtype, ptr, shape = arr, _lib.QASR_F32, arr.ctypes.data, arr.shape
            cshape = (ctypes.c_int64 * len(shape))(*shape)
            self._check(self.lib.qasr_decoder_set_weight(self.ptr, name.encode(), ctypes.c_void_p(ptr), dtype, len(shape), cshape))
        self._check(self.lib.qasr_decoder_finalize(self.ptr))
        self._ready = True

    @property
    def embed_tokens(self) -> torch.Tensor:
        """The embedding table (vocab, hidden) as a bf16 CUDA tensor view of the library's copy (decoder.py:218);
        pass it to ``prepare_inputs``."""
        ptr, dt = ctypes.c_void_p(), ctypes.c_int()
        self._check(self.lib.qasr_decoder_embed_table(self.ptr, ctypes.byref(ptr), ctypes.byref(dt)))
        return _wrap_device_bf16(ptr.value, (self.config.vocab_size, self.config.hidden_size), self.torch_device, owner=self)

    # ------------------------------------------------------------------ forward
    def prefill(self, embeddings, seq_offsets: Optional[Sequence[int]] = None, return_cache: bool = True, all_logits: bool = False,
                return_hidden: bool = False):
        """Prefill a varlen-packed batch of prompts.

        embeddings: ``(n, hidden)`` or ``(1, n, hidden)`` CUDA tensor / DeviceArray (fp32 or bf16), the output of
        ``prepare_inputs`` (several prompts concatenated along the first axis); ``seq_offsets``: B + 1 prompt boundaries
        (default: one prompt).  Returns ``(last_logits (B, vocab) fp32, KVCache | None)`` and, when asked, the logits of
        every position ``(n, vocab)`` and/or the final residual stream ``(n, hidden)`` appended to the tuple."""
        if not self._ready:
            raise _lib.QasrError("decoder weights not loaded")
        emb = embeddings.tensor if isinstance(embeddings, DeviceArray) else embeddings
        if not isinstance(emb, torch.Tensor) or not emb.is_cuda:
            raise ValueError("embeddings must be a CUDA tensor")
        if emb.ndim == 3:
            if emb.shape[0] != 1:
                raise ValueError("pass several prompts varlen-packed as (n, hidden) with seq_offsets")
            emb = emb[0]
        c = self.config
        if emb.ndim != 2 or emb.shape[1] != c.hidden_size:
            raise ValueError(f"embeddings must have shape (n, {c.hidden_size}), got {tuple(emb.shape)}")
        dts = {torch.float32: _lib.QASR_F32, torch.bfloat16: _lib.QASR_BF16}
        if emb.dtype not in dts:
            raise ValueError("embeddings must be float32 or bfloat16")
        emb = emb.contiguous()
        n = int(emb.shape[0])
        offs = np.asarray([0, n] if seq_offsets is None else list(seq_offsets), dtype=np.int64)
        if offs.ndim != 1 or len(offs) < 2 or offs[0] != 0 or offs[-1] != n or np.any(np.diff(offs) <= 0):
            raise ValueError("seq_offsets must start at 0, increase strictly and end at the number of embedding rows")
        B = len(offs) - 1
        dev = self.torch_device
        kv = c.num_key_value_heads * c.head_dim
        with torch.cuda.device(dev):
            last = torch.empty((B, c.vocab_size), dtype=torch.float32, device=dev)
            full = torch.empty((n, c.vocab_size), dtype=torch.float32, device=dev) if all_logits else None
            hid = torch.empty((n, c.hidden_size), dtype=torch.float32, device=dev) if return_hidden else None
            keys = torch.empty((c.num_hidden_layers, n, kv), dtype=torch.bfloat16, device=dev) if return_cache else None
            vals = torch.empty((c.num_hidden_layers, n, kv), dtype=torch.bfloat16, device=dev) if return_cache else None
            p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None and t.numel() else None  # noqa: E731
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            self._check(self.lib.qasr_decoder_prefill(self.ptr, p(emb), dts[emb.dtype], runtime.i64_ptr(offs), B, p(last), p(full), p(hid),
                                                      p(keys), p(vals), stream))
        cache = None
        if return_cache:
            cache = KVCache(keys.view(c.num_hidden_layers, n, c.num_key_value_heads, c.head_dim),
                            vals.view(c.num_hidden_layers, n, c.num_key_value_heads, c.head_dim), offs)
        out = [DeviceArray(last), cache]
        if all_logits:
            out.append(DeviceArray(full))
        if return_hidden:
            out.append(DeviceArray(hid))
        return tuple(out)

    def __call__(self, inputs, cache=None, is_embeds: bool = False) -> DeviceArray:
        """Reference-shaped forward for ONE prompt (decoder.py:223-253): ``(1, T, hidden)`` embeddings (or ``(1, T)`` token
        ids) -> logits ``(1, T, vocab)``.  Only the stateless / first-call case (``cache`` empty) is on this path."""
        if cache is not None and getattr(cache, "keys", None) not in (None, []):
            raise NotImplementedError("incremental decoding (a non-empty KV cache) is outside the B200 prefill path")
        if not is_embeds:
            ids = torch.as_tensor(np.asarray(inputs), device=self.torch_device).long().reshape(-1)
            inputs = self.embed_tokens[ids]
        _, _, full = self.prefill(inputs, return_cache=False, all_logits=True)
        return DeviceArray(full.tensor.unsqueeze(0))


def _wrap_device_bf16(ptr: int, shape: Tuple[int, ...], device: torch.device, owner=None) -> torch.Tensor:
    """View a device pointer owned by libqasr as a bf16 torch tensor (no copy) through the CUDA array interface."""

    class _Ext:
        pass

    n = int(np.prod(shape))
    ext = _Ext()
    ext.__cuda_array_interface__ = {"shape": (n,), "typestr": "<u2", "data": (int(ptr), False), "version": 3}
    ext._owner = owner
    with torch.cuda.device(device):
        t = torch.as_tensor(ext, device=device)
    return t.view(torch.bfloat16).view(*shape)


def load_decoder_weights(decoder: TextDecoder, model_path) -> None:
    """Load ``model.safetensors`` from a local directory: keys with the ``model.`` prefix, prefix stripped
    (reference load_decoder_weights, decoder.py:257-291; a hub repo id is resolved with ``snapshot_download`` like there)."""
    from safetensors import safe_open

    from ._hub import model_dir

    path = model_dir(model_path)

    def items():
        with safe_open(str(path / "model.safetensors"), framework="pt") as f:
            for key in f.keys():
                if key.startswith(PREFIX):
                    yield key[len(PREFIX):], f.get_tensor(key)

    decoder.load_weights(items())
