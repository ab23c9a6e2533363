"""`mlx.core` stand-in backed by torch-CPU fp32 (TEST INFRASTRUCTURE, authoring container only).

MLX is not installable here (no wheel for this platform, no network), so the reference's modules are executed
UNMODIFIED on top of this stand-in (oracle/reference_ref.py).  Only the surface the reference touches is provided:
  audio.py:278 (`mx.array(np)`) | encoder.py:33-40,224-229,263-323 | decoder.py:56-81,141-178,246-253 |
  generate.py:46-81 | model.py:267,335,420 (`clear_cache`, `eval`).
MLX semantics restated here (MLX public API behaviour; these are the assumptions SURVEY.md §8c lists):
  * `array(...)`: python ints -> int32, python floats / float64 numpy -> float32, bool -> bool_.
  * `arange(n)` -> int32 unless a dtype is given; `full(shape, v)` / `zeros(shape)` -> float32.
  * indexing, slice-assignment (`a[s:e, s:e] = 0.0`), `a[None]`, `.at[idx].add(v)` (functional scatter-add).
  * `transpose(*axes)` is a full permutation (numpy style), `.T` reverses axes, `reshape(*shape)`.
  * binary ops follow numpy-style broadcasting; bool * python float -> float32; float32 (+) int32 -> float32.
  * `fast.scaled_dot_product_attention(q, k, v, scale=, mask=)` = softmax(scale * q k^T + mask, fp32) v, with every
    KV head shared by n_heads / n_kv_heads consecutive query heads (GQA).
  * `fast.rope` is in mlx/nn.py (RoPE module).
Everything is eager: `eval` is a no-op.
"""
from __future__ import annotations

import builtins

import numpy as np
import torch

float32 = torch.float32
float16 = torch.float16
bfloat16 = torch.bfloat16
int32 = torch.int32
int64 = torch.int64
uint32 = torch.int64  # torch has no uint32 arithmetic; only used for indices
bool_ = torch.bool

torch.set_grad_enabled(False)


def _raw(x):
    return x._t if isinstance(x, array) else x


def _to_tensor(x, dtype=None) -> torch.Tensor:
    if isinstance(x, array):
        t = x._t
    elif isinstance(x, torch.Tensor):
        t = x
    elif isinstance(x, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(x))
        if t.dtype == torch.float64:
            t = t.float()
        elif t.dtype == torch.int64 and dtype is None:
            t = t.to(torch.int32) if t.numel() == 0 or int(t.abs().max()) < 2 ** 31 else t
    elif isinstance(x, (bool, np.bool_)):
        t = torch.tensor(bool(x))
    elif isinstance(x, (int, np.integer)):
        t = torch.tensor(int(x), dtype=torch.int32)
    elif isinstance(x, (float, np.floating)):
        t = torch.tensor(float(x), dtype=torch.float32)
    else:  # nested python lists
        a = np.asarray(x)
        return _to_tensor(a, dtype)
    if dtype is not None:
        t = t.to(dtype)
    return t


def _operand(x, other: torch.Tensor) -> torch.Tensor:
    """python scalars are weakly typed (they take the array's floating dtype; bool/int arrays become float32 for floats)."""
    if isinstance(x, array):
        return x._t
    if isinstance(x, bool):
        return torch.tensor(x)
    if isinstance(x, int):
        return torch.tensor(x, dtype=other.dtype if other.dtype != torch.bool else torch.int32)
    if isinstance(x, float):
        return torch.tensor(x, dtype=other.dtype if other.dtype.is_floating_point else torch.float32)
    return _to_tensor(x)


class _At:
    def __init__(self, owner, idx=None):
        self._owner, self._idx = owner, idx

    def __getitem__(self, idx):
        return _At(self._owner, idx)

    def add(self, value):
        out = self._owner._t.clone()
        out[_index(self._idx)] += _raw(value)
        return array(out)


def _index(idx):
    if isinstance(idx, tuple):
        return tuple(_raw(i).long() if isinstance(i, array) and _raw(i).dtype != torch.bool else _raw(i) for i in idx)
    if isinstance(idx, array):
        return idx._t.long() if idx._t.dtype != torch.bool else idx._t
    return idx


class array:
    __array_priority__ = 1000

    def __init__(self, value, dtype=None):
        self._t = _to_tensor(value, dtype)

    # -- numpy / python interop
    def __array__(self, dtype=None, copy=None):
        a = self._t.float().numpy() if self._t.dtype == torch.bfloat16 else self._t.numpy()
        return a.astype(dtype) if dtype is not None else a

    def item(self):
        return self._t.item()

    def tolist(self):
        return self._t.tolist()

    def __len__(self):
        return self._t.shape[0]

    def __repr__(self):
        return f"array({self._t})"

    # -- metadata
    @property
    def shape(self):
        return tuple(self._t.shape)

    @property
    def ndim(self):
        return self._t.ndim

    @property
    def dtype(self):
        return self._t.dtype

    @property
    def size(self):
        return self._t.numel()

    @property
    def T(self):
        return array(self._t.permute(*reversed(range(self._t.ndim))))

    @property
    def at(self):
        return _At(self)

    # -- shape ops
    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return array(self._t.reshape(*shape))

    def transpose(self, *axes):
        if len(axes) == 1 and isinstance(axes[0], (tuple, list)):
            axes = tuple(axes[0])
        if not axes:
            axes = tuple(reversed(range(self._t.ndim)))
        return array(self._t.permute(*axes))

    def astype(self, dtype):
        return array(self._t.to(dtype))

    def __getitem__(self, idx):
        idx = _index(idx)
        items = idx if isinstance(idx, tuple) else (idx,)
        if builtins.any(isinstance(i, slice) and i.step is not None and i.step < 0 for i in items):
            # torch has no negative-step slices (generate.py:139 `mx.sort(logits)[::-1]`): turn them into index lists
            dim, out = 0, []
            for i in items:
                if isinstance(i, slice) and i.step is not None and i.step < 0:
                    i = torch.tensor(list(range(self._t.shape[dim])[i]), dtype=torch.long)
                if i is not None:
                    dim += 1
                out.append(i)
            idx = tuple(out)
        return array(self._t[idx])

    def __setitem__(self, idx, value):
        self._t[_index(idx)] = _raw(value) if isinstance(value, array) else value

    # -- arithmetic
    def _bin(self, other, fn, reverse=False):
        o = _operand(other, self._t)
        a, b = (o, self._t) if reverse else (self._t, o)
        if fn is torch.matmul and a.dtype != b.dtype:
            dt = torch.promote_types(a.dtype, b.dtype)
            a, b = a.to(dt), b.to(dt)
        return array(fn(a, b))

    def __add__(self, o): return self._bin(o, torch.add)
    def __radd__(self, o): return self._bin(o, torch.add, True)
    def __sub__(self, o): return self._bin(o, torch.sub)
    def __rsub__(self, o): return self._bin(o, torch.sub, True)
    def __mul__(self, o): return self._bin(o, torch.mul)
    def __rmul__(self, o): return self._bin(o, torch.mul, True)
    def __truediv__(self, o): return self._bin(o, torch.true_divide)
    def __rtruediv__(self, o): return self._bin(o, torch.true_divide, True)
    def __matmul__(self, o): return self._bin(o, torch.matmul)
    def __lt__(self, o): return self._bin(o, torch.lt)
    def __le__(self, o): return self._bin(o, torch.le)
    def __gt__(self, o): return self._bin(o, torch.gt)
    def __ge__(self, o): return self._bin(o, torch.ge)
    def __eq__(self, o): return self._bin(o, torch.eq)  # noqa: E704
    def __neg__(self): return array(-self._t)
    def __pow__(self, o): return self._bin(o, torch.pow)
    def __ne__(self, o): return self._bin(o, torch.ne)  # noqa: E704
    def __and__(self, o): return self._bin(o, torch.logical_and)
    def __or__(self, o): return self._bin(o, torch.logical_or)
    def __invert__(self): return array(~self._t)
    def __bool__(self): return bool(self._t)
    def __int__(self): return int(self._t)
    def __float__(self): return float(self._t)
    def __iter__(self): return (array(t) for t in self._t)
    __hash__ = None

    def sum(self, axis=None, keepdims=False):
        return array(self._t.sum() if axis is None else self._t.sum(dim=axis, keepdim=keepdims))

    def max(self, axis=None, keepdims=False):
        return array(self._t.max() if axis is None else self._t.amax(dim=axis, keepdim=keepdims))

    def min(self, axis=None, keepdims=False):
        return array(self._t.min() if axis is None else self._t.amin(dim=axis, keepdim=keepdims))

    def mean(self, axis=None, keepdims=False):
        return array(self._t.mean() if axis is None else self._t.mean(dim=axis, keepdim=keepdims))

    def all(self):
        return array(self._t.all())

    def any(self):
        return array(self._t.any())


def _wrap1(fn):
    def f(x, *a, **k):
        return array(fn(_to_tensor(x), *a, **k))
    return f


exp = _wrap1(torch.exp)
sin = _wrap1(torch.sin)
cos = _wrap1(torch.cos)
sqrt = _wrap1(torch.sqrt)
rsqrt = _wrap1(torch.rsqrt)
sigmoid = _wrap1(torch.sigmoid)
erf = _wrap1(torch.erf)
abs = _wrap1(torch.abs)  # noqa: A001


def arange(*args, dtype=None):
    if dtype is None:
        dtype = torch.float32 if builtins.any(isinstance(a, float) for a in args) else torch.int32
    return array(torch.arange(*args, dtype=dtype))


def zeros(shape, dtype=float32):
    return array(torch.zeros(tuple(shape) if not isinstance(shape, int) else (shape,), dtype=dtype))


def ones(shape, dtype=float32):
    return array(torch.ones(tuple(shape) if not isinstance(shape, int) else (shape,), dtype=dtype))


def full(shape, vals, dtype=None):
    v = _to_tensor(vals, dtype)
    return array(torch.full(tuple(shape) if not isinstance(shape, int) else (shape,), v.item(), dtype=v.dtype))


def concatenate(arrays, axis=0):
    return array(torch.cat([_to_tensor(a) for a in arrays], dim=axis))


def stack(arrays, axis=0):
    return array(torch.stack([_to_tensor(a) for a in arrays], dim=axis))


def where(cond, a, b):
    c = _to_tensor(cond)
    ta = _operand(a, _to_tensor(b) if not isinstance(b, (int, float)) else torch.zeros(()))
    tb = _operand(b, ta)
    return array(torch.where(c, ta, tb))


def softmax(x, axis=-1):
    return array(torch.softmax(_to_tensor(x).float(), dim=axis).to(_to_tensor(x).dtype))


def argmax(x, axis=None):
    t = _to_tensor(x)
    return array(t.argmax() if axis is None else t.argmax(dim=axis))


def matmul(a, b):
    return array(torch.matmul(_to_tensor(a), _to_tensor(b)))


def mean(x, axis=None, keepdims=False):
    return array(x).mean(axis, keepdims)


def sum(x, axis=None, keepdims=False):  # noqa: A001
    return array(x).sum(axis, keepdims)


def max(x, axis=None, keepdims=False):  # noqa: A001
    return array(x).max(axis, keepdims)


def sort(x, axis=-1):
    return array(torch.sort(_to_tensor(x), dim=axis).values)


def argsort(x, axis=-1):
    return array(torch.argsort(_to_tensor(x), dim=axis, stable=True).to(torch.int32))


def cumsum(x, axis=None):
    t = _to_tensor(x)
    return array(torch.cumsum(t.flatten() if axis is None else t, dim=0 if axis is None else axis))


def all(x):  # noqa: A001
    return array(_to_tensor(x).all())


def any(x):  # noqa: A001
    return array(_to_tensor(x).any())


def array_equal(a, b):
    return array(torch.equal(_to_tensor(a), _to_tensor(b)))


def allclose(a, b, rtol=1e-5, atol=1e-8):
    return array(torch.allclose(_to_tensor(a).float(), _to_tensor(b).float(), rtol=rtol, atol=atol))


def isnan(x):
    return array(torch.isnan(_to_tensor(x)))


def isinf(x):
    return array(torch.isinf(_to_tensor(x)))


class _Random:
    @staticmethod
    def seed(s):
        torch.manual_seed(int(s))

    @staticmethod
    def categorical(logits, axis=-1):
        p = torch.softmax(_to_tensor(logits).float(), dim=axis)
        return array(torch.multinomial(p.reshape(-1, p.shape[-1]), 1).reshape(p.shape[:-1]).to(torch.int32))

    @staticmethod
    def normal(shape=(), dtype=float32):
        return array(torch.randn(tuple(shape), dtype=dtype))

    @staticmethod
    def uniform(low=0.0, high=1.0, shape=(), dtype=float32):
        return array(torch.rand(tuple(shape), dtype=dtype) * (high - low) + low)


random = _Random()


def eval(*args):  # noqa: A001 - MLX's graph evaluation; this stand-in is eager
    return None


def clear_cache():
    return None


def load(path):
    """`mx.load("model.safetensors")` -> dict[str, array] (dtypes kept, bf16 included)."""
    from safetensors.torch import load_file

    return {k: array(v) for k, v in load_file(str(path)).items()}


class _Fast:
    @staticmethod
    def scaled_dot_product_attention(q, k, v, *, scale, mask=None):
        tq, tk, tv = _to_tensor(q), _to_tensor(k), _to_tensor(v)
        rep = tq.shape[1] // tk.shape[1]
        if rep > 1:
            tk, tv = tk.repeat_interleave(rep, dim=1), tv.repeat_interleave(rep, dim=1)
        s = torch.matmul(tq.float(), tk.float().transpose(-1, -2)) * scale
        if mask is not None:
            m = _to_tensor(mask)
            s = torch.where(m, s, torch.full_like(s, float("-inf"))) if m.dtype == torch.bool else s + m.float()
        return array(torch.matmul(torch.softmax(s, dim=-1), tv.float()).to(tq.dtype))

    @staticmethod
    def rope(x, dims, *, traditional, base, scale, offset):
        t = _to_tensor(x)
        T = t.shape[-2]
        half = dims // 2
        inv = float(base) ** (-torch.arange(half, dtype=torch.float32) / half)
        ang = ((torch.arange(T, dtype=torch.float32) + float(offset)) * scale)[:, None] * inv[None, :]
        c, s = torch.cos(ang), torch.sin(ang)
        rot, rest = t[..., :dims].float(), t[..., dims:]
        if traditional:  # interleaved pairs (2i, 2i+1)
            x1, x2 = rot[..., 0::2], rot[..., 1::2]
            out = torch.stack([x1 * c - x2 * s, x1 * s + x2 * c], dim=-1).flatten(-2)
        else:            # pairs (i, i + dims/2)
            x1, x2 = rot[..., :half], rot[..., half:]
            out = torch.cat([x1 * c - x2 * s, x1 * s + x2 * c], dim=-1)
        return array(torch.cat([out.to(t.dtype), rest], dim=-1))


fast = _Fast()
