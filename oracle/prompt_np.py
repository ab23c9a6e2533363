"""numpy restatement of the reference's prompt assembly (TEST INFRASTRUCTURE).

prepare_inputs: /root/reference/src/qwen3_asr_mlx/generate.py:20-81 — embed every id, then overwrite the
audio-pad positions, in order, with the encoder rows cast to the embedding dtype; ValueError when the counts differ;
plain embeddings when there is no pad.  build_prompt: tokenizer.py:16-86 (ids restated from the reference).
PINNED: tests/golden/prompt_reference.npz holds the reference's own build_prompt ids and prepare_inputs outputs
(oracle/reference_ref.py); tests/test_reference_pin.py requires exact equality.
"""
import numpy as np

AUDIO_PAD = 151676
PREFIX = [151644, 8948, 198, 151645, 198, 151644, 872, 198, 151669]  # tokenizer.py:27-37
SUFFIX = [151670, 151645, 198, 151644, 77091, 198]                    # tokenizer.py:39-46


def build_prompt(n_audio_tokens, language_name_tokens=None):
    return PREFIX + [AUDIO_PAD] * n_audio_tokens + SUFFIX + [11528] + list(language_name_tokens or []) + [151704]


def prepare_inputs(encoder_output, input_ids, table, audio_pad_id=AUDIO_PAD):
    emb = table[np.asarray(input_ids)].copy()
    pos = [i for i, t in enumerate(input_ids) if t == audio_pad_id]
    enc = np.asarray(encoder_output).reshape(-1, table.shape[1])
    if not pos:
        return emb[None]
    if len(pos) != enc.shape[0]:
        raise ValueError(f"Number of audio-pad tokens ({len(pos)}) does not match encoder output length ({enc.shape[0]}).")
    emb[pos] = enc.astype(emb.dtype)
    return emb[None]
