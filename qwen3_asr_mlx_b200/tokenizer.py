"""Prompt layout of Qwen3-ASR (reference src/qwen3_asr_mlx/tokenizer.py:16-86).

Only what the audio-encoding path's consumer needs: the special token ids of the Qwen3-ASR vocabulary
and ``build_prompt``.  BPE encoding / decoding of text stays with the decoder side (out of scope).
"""
from __future__ import annotations

from typing import List, Optional

# Qwen3-ASR vocabulary facts (tokenizer.py:16-49)
AUDIO_START_TOKEN_ID = 151669
AUDIO_END_TOKEN_ID = 151670
AUDIO_PAD_TOKEN_ID = 151676
IM_START_TOKEN_ID = 151644
IM_END_TOKEN_ID = 151645
ENDOFTEXT_TOKEN_ID = 151643
ASR_TEXT_TOKEN_ID = 151704
EOS_TOKEN_IDS = frozenset({ENDOFTEXT_TOKEN_ID, IM_END_TOKEN_ID})

_SYSTEM, _USER, _ASSISTANT, _NEWLINE, _LANGUAGE = 8948, 872, 77091, 198, 11528


def build_prompt(n_audio_tokens: int, language_name_tokens: Optional[List[int]] = None) -> List[int]:
    """input_ids of an inference prompt (tokenizer.py:56-86):

    ``<|im_start|>system\\n<|im_end|>\\n<|im_start|>user\\n<|audio_start|>`` + N x ``<|audio_pad|>`` +
    ``<|audio_end|><|im_end|>\\n<|im_start|>assistant\\n`` + ``language`` + name tokens + ``<asr_text>``.
    """
    head = [IM_START_TOKEN_ID, _SYSTEM, _NEWLINE, IM_END_TOKEN_ID, _NEWLINE, IM_START_TOKEN_ID, _USER, _NEWLINE, AUDIO_START_TOKEN_ID]
    tail = [AUDIO_END_TOKEN_ID, IM_END_TOKEN_ID, _NEWLINE, IM_START_TOKEN_ID, _ASSISTANT, _NEWLINE]
    return head + [AUDIO_PAD_TOKEN_ID] * int(n_audio_tokens) + tail + [_LANGUAGE] + list(language_name_tokens or []) + [ASR_TEXT_TOKEN_ID]
