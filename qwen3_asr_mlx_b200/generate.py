"""`prepare_inputs`: the consumer of the audio-encoding path (reference src/qwen3_asr_mlx/generate.py:20-81).

Text-token rows come from the decoder's embedding table, ``<|audio_pad|>`` rows from the encoder output, in one
gather kernel (``qasr_prepare_inputs``) instead of the reference's per-token ``.at[].add`` loop.  Sampling and the
token loop of the reference's ``generate`` stay out of scope.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib, runtime
from ._array import DeviceArray
from .tokenizer import AUDIO_PAD_TOKEN_ID


def prepare_inputs(encoder_output, input_ids: Sequence[int], embed_tokens, audio_pad_id: int = AUDIO_PAD_TOKEN_ID) -> DeviceArray:
    """Replace audio-pad token embeddings with encoder output features.

    encoder_output: ``(1, n_audio, hidden)`` or ``(n_audio, hidden)`` device array (fp32 or bf16);
    embed_tokens: the decoder's embedding table ``(vocab, hidden)`` as a CUDA tensor / DeviceArray (fp32 or bf16);
    returns ``(1, len(input_ids), hidden)`` in the table's dtype.  A pad count that differs from the number of
    encoder rows raises ``ValueError`` (generate.py:58-62); a prompt without pads returns the text embeddings.
    """
    table = embed_tokens.tensor if isinstance(embed_tokens, DeviceArray) else embed_tokens
    if not isinstance(table, torch.Tensor) or not table.is_cuda or table.ndim != 2:
        raise ValueError("embed_tokens must be a (vocab, hidden) CUDA tensor")
    audio = encoder_output.tensor if isinstance(encoder_output, DeviceArray) else encoder_output
    audio = audio.reshape(-1, audio.shape[-1]).contiguous()
    h = runtime.frontend_handle(table.device.index)
    ids = np.ascontiguousarray(np.asarray(list(input_ids), dtype=np.int32))
    dt = {torch.float32: _lib.QASR_F32, torch.bfloat16: _lib.QASR_BF16}
    if table.dtype not in dt or audio.dtype not in dt:
        raise ValueError("embedding table and encoder output must be float32 or bfloat16")
    table = table.contiguous()
    out = torch.empty((len(ids), table.shape[1]), dtype=table.dtype, device=table.device)
    with torch.cuda.device(table.device):
        h.check(h.lib.qasr_prepare_inputs(h.ptr, ids.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), len(ids), ctypes.c_void_p(table.data_ptr()),
                                          dt[table.dtype], table.shape[0], table.shape[1], ctypes.c_void_p(audio.data_ptr()), dt[audio.dtype],
                                          audio.shape[0], int(audio_pad_id), ctypes.c_void_p(out.data_ptr()), h.stream_ptr()))
    return DeviceArray(out.unsqueeze(0))
