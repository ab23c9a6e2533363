"""DeviceArray: the array type the B200 path hands back where the reference returns ``mx.array``.

It wraps a CUDA ``torch.Tensor`` (torch is used for device memory and streams only) and offers
what the reference's callers use: ``.shape``, ``.ndim``, ``.dtype``, indexing, ``np.array(x)``
(device-to-host copy) and the DLPack protocol.
"""
from __future__ import annotations

import numpy as np
import torch


class DeviceArray:
    __slots__ = ("tensor",)

    def __init__(self, tensor: torch.Tensor):
        self.tensor = tensor

    @property
    def shape(self):
        return tuple(self.tensor.shape)

    @property
    def ndim(self) -> int:
        return self.tensor.ndim

    @property
    def dtype(self):
        return self.tensor.dtype

    def __len__(self) -> int:
        return self.tensor.shape[0]

    def __getitem__(self, idx):
        return DeviceArray(self.tensor[idx])

    def numpy(self) -> np.ndarray:
        t = self.tensor
        if t.dtype == torch.bfloat16:
            t = t.float()
        return t.detach().cpu().numpy()

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a.astype(dtype) if dtype is not None else a

    def __dlpack__(self, *args, **kwargs):
        return self.tensor.__dlpack__(*args, **kwargs)

    def __dlpack_device__(self):
        return self.tensor.__dlpack_device__()

    def __repr__(self) -> str:
        return f"DeviceArray(shape={self.shape}, dtype={self.dtype}, device={self.tensor.device})"


def as_device_f32(x, device: torch.device) -> torch.Tensor:
    """numpy / torch / DeviceArray / DLPack producer -> contiguous float32 tensor on ``device``."""
    if isinstance(x, DeviceArray):
        t = x.tensor
    elif isinstance(x, torch.Tensor):
        t = x
    elif isinstance(x, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    elif hasattr(x, "__dlpack__"):
        t = torch.from_dlpack(x)
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float32)))
    return t.to(device=device, dtype=torch.float32, non_blocking=True).contiguous()
