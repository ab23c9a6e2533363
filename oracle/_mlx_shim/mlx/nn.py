"""`mlx.nn` stand-in backed by torch-CPU fp32 (TEST INFRASTRUCTURE, authoring container only; see mlx/core.py).

Modules the reference constructs (encoder.py:60-63,98-104,148-191; decoder.py:93-95,122-128,189-192,219-221) with the
semantics of MLX's public API:
  * Linear(in, out, bias=True): weight (out, in), y = x W^T + b.
  * Conv2d(in, out, kernel_size, stride, padding): NHWC input, weight (out, kH, kW, in), cross-correlation, zero padding.
  * LayerNorm(dims, eps=1e-5, affine=True): biased variance over the last axis.
  * RMSNorm(dims, eps): x * rsqrt(mean(x^2) + eps) * weight.
  * Embedding(n, dims): weight (n, dims), lookup by integer index.
  * RoPE(dims, traditional=False, base): mx.fast.rope over the second-to-last axis, positions offset + t.
  * gelu = exact erf form; silu = x * sigmoid(x).
  * Module.parameters() walks public attributes (names starting with "_" are not parameters: encoder.py:38 keeps the
    positional table in `_positional_embedding`); load_weights(list[(dotted name, array)], strict=True) raises
    ValueError on unknown names, missing names or shape mismatches, like MLX.
Freshly constructed parameters use MLX's default distributions; every oracle run overwrites them with load_weights.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import core as mx


class Module:
    def __call__(self, *args, **kwargs):
        raise NotImplementedError

    # -- parameter tree
    def _children(self):
        for k, v in vars(self).items():
            if k.startswith("_"):
                continue
            yield k, v

    def parameters(self):
        def walk(v):
            if isinstance(v, mx.array):
                return v
            if isinstance(v, Module):
                return v.parameters()
            if isinstance(v, (list, tuple)):
                out = [walk(i) for i in v]
                return out if any(o is not None for o in out) else None
            if isinstance(v, dict):
                return {k: walk(i) for k, i in v.items()}
            return None

        return {k: w for k, w in ((k, walk(v)) for k, v in self._children()) if w is not None}

    def _flat(self, prefix=""):
        def walk(v, name):
            if isinstance(v, mx.array):
                yield name, v
            elif isinstance(v, dict):
                for k, i in v.items():
                    yield from walk(i, f"{name}.{k}")
            elif isinstance(v, list):
                for k, i in enumerate(v):
                    if i is not None:
                        yield from walk(i, f"{name}.{k}")

        for k, v in self.parameters().items():
            yield from walk(v, prefix + k)

    def load_weights(self, weights, strict: bool = True):
        if isinstance(weights, dict):
            weights = list(weights.items())
        have = dict(self._flat())
        given = dict(weights)
        if strict:
            extra = sorted(set(given) - set(have))
            if extra:
                raise ValueError(f"Received parameters not in model: {' '.join(extra)}.")
            missing = sorted(set(have) - set(given))
            if missing:
                raise ValueError(f"Missing parameters: {' '.join(missing)}.")
        for name, value in given.items():
            if name not in have:
                continue
            value = value if isinstance(value, mx.array) else mx.array(value)
            if tuple(value.shape) != tuple(have[name].shape):
                raise ValueError(f"Expected shape {have[name].shape} but received shape {value.shape} for parameter {name}")
            obj = self
            *path, leaf = name.split(".")
            for p in path:
                obj = obj[int(p)] if isinstance(obj, (list, tuple)) else (obj[p] if isinstance(obj, dict) else getattr(obj, p))
            setattr(obj, leaf, value)
        return self

    def eval(self):
        return self

    def freeze(self):
        return self


def _promote(x, *params):
    """MLX type promotion: an fp32 activation meeting bf16 / fp16 parameters computes (and returns) fp32 -- this is why the
    reference's activations are fp32 even with the bf16 checkpoint."""
    t = mx._to_tensor(x)
    dt = t.dtype
    for p_ in params:
        if p_ is not None:
            dt = torch.promote_types(dt, p_.dtype)
    return (t.to(dt),) + tuple(None if p_ is None else p_.to(dt) for p_ in params)


def _uniform(shape, scale):
    return mx.array((torch.rand(shape) * 2.0 - 1.0) * scale)


class Linear(Module):
    def __init__(self, input_dims: int, output_dims: int, bias: bool = True):
        s = 1.0 / math.sqrt(input_dims)
        self.weight = _uniform((output_dims, input_dims), s)
        if bias:
            self.bias = _uniform((output_dims,), s)

    def __call__(self, x):
        b = getattr(self, "bias", None)
        t, w, bt = _promote(x, self.weight._t, None if b is None else b._t)
        return mx.array(F.linear(t, w, bt))


class Conv2d(Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True):
        kh, kw = (kernel_size, kernel_size) if isinstance(kernel_size, int) else kernel_size
        s = 1.0 / math.sqrt(in_channels * kh * kw)
        self.weight = _uniform((out_channels, kh, kw, in_channels), s)
        if bias:
            self.bias = mx.zeros((out_channels,))
        self._stride, self._padding, self._dilation, self._groups = stride, padding, dilation, groups

    def __call__(self, x):
        b = getattr(self, "bias", None)
        t, w, bt = _promote(x, self.weight._t, None if b is None else b._t)
        t = t.permute(0, 3, 1, 2)                                       # NHWC -> NCHW
        w = w.permute(0, 3, 1, 2)                                       # (O,kH,kW,I) -> (O,I,kH,kW)
        y = F.conv2d(t, w, bt, stride=self._stride, padding=self._padding,
                     dilation=self._dilation, groups=self._groups)
        return mx.array(y.permute(0, 2, 3, 1))                           # back to NHWC


class LayerNorm(Module):
    def __init__(self, dims: int, eps: float = 1e-5, affine: bool = True, bias: bool = True):
        if affine:
            self.weight = mx.ones((dims,))
            if bias:
                self.bias = mx.zeros((dims,))
        self._eps, self._dims = eps, dims

    def __call__(self, x):
        w, b = getattr(self, "weight", None), getattr(self, "bias", None)
        t, wt, bt = _promote(x, None if w is None else w._t, None if b is None else b._t)
        return mx.array(F.layer_norm(t, (self._dims,), wt, bt, self._eps))


class RMSNorm(Module):
    def __init__(self, dims: int, eps: float = 1e-5):
        self.weight = mx.ones((dims,))
        self._eps = eps

    def __call__(self, x):
        t, w = _promote(x, self.weight._t)
        tf = t.float()
        return mx.array((tf * torch.rsqrt(tf.pow(2).mean(-1, keepdim=True) + self._eps)).to(t.dtype) * w)


class Embedding(Module):
    def __init__(self, num_embeddings: int, dims: int):
        self.weight = mx.array(torch.randn(num_embeddings, dims) * math.sqrt(1.0 / dims))

    def __call__(self, x):
        # MLX gathers without bounds checks (the reference's tests look up id 151676 in a 512-row table): clamp, never raise
        return mx.array(self.weight._t[mx._to_tensor(x).long().clamp(0, self.weight._t.shape[0] - 1)])

    def as_linear(self, x):
        return mx.array(mx._to_tensor(x) @ self.weight._t.T)


class RoPE(Module):
    def __init__(self, dims: int, traditional: bool = False, base: float = 10000.0, scale: float = 1.0):
        self._dims, self._traditional, self._base, self._scale = dims, traditional, base, scale

    def __call__(self, x, offset: int = 0):
        return mx.fast.rope(x, self._dims, traditional=self._traditional, base=self._base, scale=self._scale, offset=offset)


def gelu(x):
    return mx.array(F.gelu(mx._to_tensor(x)))  # exact erf form


def silu(x):
    return mx.array(F.silu(mx._to_tensor(x)))


def relu(x):
    return mx.array(F.relu(mx._to_tensor(x)))
