"""CPU: the two independent encoder restatements agree, and honour the reference's shape
contract (reference tests/test_encoder.py:41-198).  The reference holds no numeric vectors for
the encoder; numeric parity is pinned by tests/test_reference_pin.py, which compares these restatements with
the reference's own encoder.py executed verbatim."""
import os

import numpy as np
import pytest

from oracle import encoder_np, encoder_torch, mel_np
from qwen3_asr_mlx_b200 import weights
from qwen3_asr_mlx_b200.config import AudioEncoderConfig
from helpers import rel_err, synth

SMALL = AudioEncoderConfig(d_model=256, encoder_layers=2, encoder_attention_heads=4, encoder_ffn_dim=512, output_dim=256)


@pytest.fixture(scope="module")
def small_params():
    return weights.random_init(SMALL, seed=7, exercise_all=True)


@pytest.mark.parametrize("n_samples", [16000, 16000 * 2 + 8000, 16000 * 9 + 4321])
def test_torch_and_numpy_restatements_agree(small_params, n_samples):
    mel = mel_np.log_mel_spectrogram_fast(synth(np.random.default_rng(n_samples), n_samples))
    a = encoder_torch.encoder_forward(small_params, SMALL, mel)
    b = encoder_np.encoder_forward(small_params, SMALL, mel)
    assert a.shape == b.shape
    assert rel_err(a, b) <= 1e-5


def test_golden_anchor(small_params, golden_dir):
    g = np.load(os.path.join(golden_dir, "encoder_small.npz"))
    mel = mel_np.log_mel_spectrogram(g["audio"])
    out = encoder_torch.encoder_forward(small_params, SMALL, mel)
    assert out.shape == g["emb"].shape == (121, 256)  # 9 full chunks (117) + f3(27) = 4
    assert rel_err(out, g["emb"]) <= 1e-5


@pytest.mark.parametrize("T,tokens", [(100, 13), (300, 39), (250, 33), (50, 7), (1, 1), (101, 14)])
def test_output_token_counts(small_params, T, tokens):
    # reference tests/test_encoder.py:64-89,161-164
    mel = np.random.default_rng(0).standard_normal((128, T)).astype(np.float32)
    out = encoder_torch.encoder_forward(small_params, SMALL, mel)
    assert out.shape == (tokens, SMALL.output_dim)
    assert np.isfinite(out).all()


def test_batched_input_drops_extra_entries(small_params):
    # reference encoder.py:249-250 and tests/test_encoder.py:91-97
    mel = np.random.default_rng(1).standard_normal((2, 128, 100)).astype(np.float32)
    a = encoder_torch.encoder_forward(small_params, SMALL, mel)
    b = encoder_torch.encoder_forward(small_params, SMALL, mel[0])
    assert np.array_equal(a, b)


def test_conv_length_rule():
    assert encoder_torch.conv_output_length(100) == 13 and encoder_torch.conv_output_length(50) == 7


def test_positional_table():
    # reference tests/test_encoder.py:171-198
    pe = encoder_torch.positional_table(13, 1024).numpy()
    assert pe.shape == (13, 1024)
    assert np.array_equal(pe, encoder_torch.positional_table(13, 1024).numpy())
    assert np.array_equal(encoder_torch.positional_table(20, 1024).numpy()[:13], pe)
    assert np.allclose(pe[0, :512], 0.0) and np.allclose(pe[0, 512:], 1.0)
    assert np.allclose(pe, encoder_np.positional_table(13, 1024), atol=1e-6)


def test_windows_are_independent(small_params):
    """Block-diagonal attention: perturbing audio of the second 8-s window leaves the first window's
    tokens unchanged (equivalent to the reference's -1e9 block mask)."""
    rng = np.random.default_rng(3)
    mel = rng.standard_normal((128, 1000)).astype(np.float32)
    mel2 = mel.copy()
    mel2[:, 800:] += 1.0
    a = encoder_torch.encoder_forward(small_params, SMALL, mel)
    b = encoder_torch.encoder_forward(small_params, SMALL, mel2)
    assert a.shape == (130, 256)
    assert np.array_equal(a[:104], b[:104]) and not np.allclose(a[104:], b[104:])
