"""CPU: the encoder oracles agree with the model authors' own PyTorch implementation of the same
audio tower (transformers' Qwen3OmniMoeAudioEncoder, see tests/upstream_hf.py).  This is the
strongest pin available offline for the encoder half: the reference's tests hold shapes only and
MLX cannot be installed here."""
import numpy as np
import pytest
import torch

from oracle import encoder_np, encoder_torch
from qwen3_asr_mlx_b200 import weights
from qwen3_asr_mlx_b200.config import AudioEncoderConfig
from qwen3_asr_mlx_b200.launcher import tokens_for_samples
from helpers import rel_err

upstream_hf = pytest.importorskip("upstream_hf")
pytest.importorskip("transformers.models.qwen3_omni_moe.modeling_qwen3_omni_moe")

def token_count(T):
    return tokens_for_samples(T * 160)


SMALL = AudioEncoderConfig(d_model=256, encoder_layers=2, encoder_attention_heads=4, encoder_ffn_dim=512, output_dim=256)


@pytest.fixture(scope="module")
def small():
    params = weights.random_init(SMALL, seed=7, exercise_all=True)
    return params, upstream_hf.build_upstream(SMALL, params)


def test_parameter_names_and_shapes_match_upstream(small):
    # strict load_state_dict inside build_upstream already raised on any mismatch; check the inventory size too
    params, model = small
    assert set(params) == set(model.state_dict())


# T >= 100 so that the tail chunk is padded to 100 frames by upstream as well (see upstream_hf docstring)
@pytest.mark.parametrize("T", [100, 250, 750, 800, 801, 1050, 1699, 3000])
def test_oracles_match_upstream(small, T):
    params, model = small
    mel = np.random.default_rng(T).standard_normal((128, T)).astype(np.float32)
    up = upstream_hf.upstream_forward(model, mel)
    a = encoder_torch.encoder_forward(params, SMALL, mel)
    assert up.shape == a.shape == (token_count(T), SMALL.output_dim)
    assert rel_err(a, up) <= 1e-5
    if T <= 1050:
        assert rel_err(encoder_np.encoder_forward(params, SMALL, mel), up) <= 1e-5


def test_token_count_rule_matches_upstream():
    from transformers.models.qwen3_omni_moe.modeling_qwen3_omni_moe import _get_feat_extract_output_lengths

    T = torch.arange(1, 6001)
    ours = torch.tensor([token_count(int(t)) for t in T])
    assert torch.equal(_get_feat_extract_output_lengths(T), ours)


def test_positional_table_matches_upstream():
    from transformers.models.qwen3_omni_moe.modeling_qwen3_omni_moe import SinusoidsPositionEmbedding

    up = SinusoidsPositionEmbedding(1500, 1024).positional_embedding[:13].numpy()
    assert np.allclose(encoder_torch.positional_table(13, 1024).numpy(), up, atol=1e-6)


def test_full_width_layers_match_upstream():
    """1.7B widths (d_model 1024, 16 heads, ffn 4096, out 2048) with 2 layers: 10.5 s, two windows."""
    cfg = AudioEncoderConfig(encoder_layers=2)
    params = weights.random_init(cfg, seed=11, exercise_all=True)
    model = upstream_hf.build_upstream(cfg, params)
    mel = np.random.default_rng(5).standard_normal((128, 1050)).astype(np.float32)
    assert rel_err(encoder_torch.encoder_forward(params, cfg, mel), upstream_hf.upstream_forward(model, mel)) <= 1e-5
