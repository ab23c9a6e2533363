"""Prompt assembly: build_prompt ids (CPU) and prepare_inputs scatter (GPU)."""
import numpy as np
import pytest

from oracle import prompt_np
from qwen3_asr_mlx_b200 import tokenizer


def test_build_prompt_ids_match_reference_layout():
    # reference tests/test_tokenizer.py:48-73: prefix / pads / suffix / language / <asr_text>
    ids = tokenizer.build_prompt(5, [6364])
    assert ids == prompt_np.build_prompt(5, [6364])
    assert ids[:9] == [151644, 8948, 198, 151645, 198, 151644, 872, 198, 151669]
    assert ids[9:14] == [tokenizer.AUDIO_PAD_TOKEN_ID] * 5
    assert ids[14:20] == [151670, 151645, 198, 151644, 77091, 198]
    assert ids[20] == 11528 and ids[21] == 6364 and ids[-1] == tokenizer.ASR_TEXT_TOKEN_ID
    assert tokenizer.build_prompt(0) == prompt_np.build_prompt(0) and len(tokenizer.build_prompt(0)) == 9 + 6 + 2
    assert tokenizer.EOS_TOKEN_IDS == frozenset({151643, 151645})


@pytest.mark.gpu
@pytest.mark.parametrize("table_dtype,audio_dtype", [("float32", "float32"), ("bfloat16", "float32"), ("bfloat16", "bfloat16")])
def test_prepare_inputs_matches_oracle(table_dtype, audio_dtype):
    import torch

    from qwen3_asr_mlx_b200 import prepare_inputs

    rng = np.random.default_rng(0)
    vocab, hidden, n_audio = 152000, 256, 37
    table = torch.from_numpy(rng.standard_normal((vocab, hidden)).astype(np.float32)).to(getattr(torch, table_dtype)).cuda()
    audio = torch.from_numpy(rng.standard_normal((1, n_audio, hidden)).astype(np.float32)).to(getattr(torch, audio_dtype)).cuda()
    ids = tokenizer.build_prompt(n_audio, [6364, 100])
    got = prepare_inputs(audio, ids, table)
    assert got.shape == (1, len(ids), hidden)
    ref = prompt_np.prepare_inputs(audio.float().cpu().numpy(), ids, table.float().cpu().numpy())
    want = torch.from_numpy(ref).to(getattr(torch, table_dtype)).float().numpy()  # audio rows are cast to the table dtype
    assert np.array_equal(got.tensor.float().cpu().numpy(), want)
    # no pads -> plain text embeddings; count mismatch -> ValueError (generate.py:55-62)
    plain = prepare_inputs(audio, [1, 2, 3], table)
    assert np.array_equal(plain.tensor.float().cpu().numpy()[0], table[[1, 2, 3]].float().cpu().numpy())
    with pytest.raises(ValueError):
        prepare_inputs(audio, tokenizer.build_prompt(n_audio - 1), table)
