"""CPU: the decoder-prefill oracle (oracle/decoder_torch.py, a restatement of reference decoder.py) agrees with the model
authors' Qwen3 implementation in transformers (tests/upstream_hf.py): logits at every position, cached keys (post q/k-norm
and RoPE) and values.  The reference's tests hold shapes only for the decoder (tests/test_decoder.py) and MLX cannot run here."""
import numpy as np
import pytest
import torch

from oracle import decoder_torch
from qwen3_asr_mlx_b200 import decoder as dec
from qwen3_asr_mlx_b200.config import TextDecoderConfig
from helpers import rel_err

upstream_hf = pytest.importorskip("upstream_hf")
pytest.importorskip("transformers.models.qwen3.modeling_qwen3")

SMALL = TextDecoderConfig(hidden_size=256, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, intermediate_size=512, vocab_size=1024)


@pytest.fixture(scope="module")
def small():
    params = dec.random_init(SMALL, seed=5, exercise_all=True)
    return params, upstream_hf.build_upstream_decoder(SMALL, params)


def test_parameter_inventory_matches_upstream(small):
    params, model = small
    assert {"model." + k for k in params} | {"lm_head.weight"} == set(model.state_dict())
    assert [n for n, _ in dec.parameter_shapes(SMALL)] == list(params)


@pytest.mark.parametrize("T", [1, 7, 64, 65, 200, 407])
def test_oracle_matches_upstream(small, T):
    params, model = small
    emb = torch.randn(T, SMALL.hidden_size, generator=torch.Generator().manual_seed(T))
    o = decoder_torch.decoder_prefill(params, SMALL, emb)
    logits, keys, values = upstream_hf.upstream_decoder_forward(model, emb)
    assert o["logits"].shape == (T, SMALL.vocab_size) and o["keys"].shape == (2, 2, T, 128)
    assert rel_err(o["logits"], logits) <= 1e-5
    assert rel_err(o["keys"], keys) <= 1e-5 and rel_err(o["values"], values) <= 1e-5


def test_full_width_layer_matches_upstream():
    """1.7B widths (hidden 2048, 16/8 heads x 128, intermediate 6144) with one layer and a reduced vocabulary."""
    cfg = TextDecoderConfig(num_hidden_layers=1, vocab_size=2048)
    params = dec.random_init(cfg, seed=9, exercise_all=True)
    model = upstream_hf.build_upstream_decoder(cfg, params)
    emb = torch.randn(90, cfg.hidden_size, generator=torch.Generator().manual_seed(3))
    o = decoder_torch.decoder_prefill(params, cfg, emb)
    logits, keys, values = upstream_hf.upstream_decoder_forward(model, emb)
    assert rel_err(o["logits"], logits) <= 1e-5 and rel_err(o["keys"], keys) <= 1e-5 and rel_err(o["values"], values) <= 1e-5


def test_causality_and_position_restart():
    params = dec.random_init(SMALL, seed=5)
    g = torch.Generator().manual_seed(0)
    a, b = torch.randn(30, 256, generator=g), torch.randn(50, 256, generator=g)
    oa = decoder_torch.decoder_prefill(params, SMALL, a)
    a2 = a.clone()
    a2[20:] += 1.0
    oa2 = decoder_torch.decoder_prefill(params, SMALL, a2)
    assert np.array_equal(oa["logits"][:20], oa2["logits"][:20]) and not np.allclose(oa["logits"][20:], oa2["logits"][20:])
    both = decoder_torch.decoder_prefill_batch(params, SMALL, torch.cat([a, b]), [0, 30, 80])
    assert np.array_equal(both[0]["logits"], oa["logits"])  # a batch is a loop over prompts, positions restart at 0
    assert np.array_equal(both[1]["logits"], decoder_torch.decoder_prefill(params, SMALL, b)["logits"])
