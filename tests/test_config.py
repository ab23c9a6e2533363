"""CPU: configuration dataclasses mirror the reference's (reference tests/test_config.py, same sample config.json)."""
import json

from qwen3_asr_mlx_b200.config import AudioEncoderConfig, ModelConfig, TextDecoderConfig

SAMPLE = {
    "audio_token_id": 151676, "audio_start_token_id": 151669, "audio_end_token_id": 151670,
    "audio_encoder_config": {"d_model": 1024, "encoder_layers": 24, "encoder_attention_heads": 16, "encoder_ffn_dim": 4096,
                             "num_mel_bins": 128, "max_source_positions": 1500, "output_dim": 2048, "n_window": 50,
                             "n_window_infer": 800, "conv_chunksize": 500, "activation_function": "gelu", "downsample_hidden_size": 480},
    "hidden_size": 2048, "num_hidden_layers": 28, "num_attention_heads": 16, "num_key_value_heads": 8, "head_dim": 128,
    "intermediate_size": 6144, "hidden_act": "silu", "vocab_size": 151936, "max_position_embeddings": 65536,
    "rms_norm_eps": 1e-6, "rope_theta": 1000000.0, "mrope_section": [24, 20, 20], "rope_interleaved": True,
}


def test_audio_encoder_config_from_dict_and_defaults():  # reference TestAudioEncoderConfig
    cfg = AudioEncoderConfig.from_dict(SAMPLE)
    assert (cfg.d_model, cfg.encoder_layers, cfg.encoder_attention_heads, cfg.encoder_ffn_dim) == (1024, 24, 16, 4096)
    assert (cfg.num_mel_bins, cfg.max_source_positions, cfg.output_dim, cfg.n_window, cfg.n_window_infer) == (128, 1500, 2048, 50, 800)
    assert (cfg.conv_chunksize, cfg.activation_function, cfg.downsample_hidden_size) == (500, "gelu", 480)
    assert AudioEncoderConfig().num_mel_bins == 128 and AudioEncoderConfig().encoder_layers == 24
    # the encoder sub-dict wins over the decoder's top-level num_hidden_layers (28)
    assert AudioEncoderConfig.from_dict({**SAMPLE, "audio_encoder_config": {"d_model": 256}}).encoder_layers == 24


def test_text_decoder_config_from_dict():  # reference TestTextDecoderConfig
    cfg = TextDecoderConfig.from_dict(SAMPLE)
    assert (cfg.hidden_size, cfg.num_hidden_layers, cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim) == (2048, 28, 16, 8, 128)
    assert (cfg.intermediate_size, cfg.hidden_act, cfg.vocab_size, cfg.max_position_embeddings) == (6144, "silu", 151936, 65536)
    assert abs(cfg.rms_norm_eps - 1e-6) < 1e-12 and cfg.rope_theta == 1_000_000.0
    assert cfg.mrope_section == [24, 20, 20] and cfg.rope_interleaved is True
    assert TextDecoderConfig.from_dict({}) == TextDecoderConfig()


def test_model_config(tmp_path):  # reference TestModelConfig
    cfg = ModelConfig.from_dict(SAMPLE)
    assert (cfg.audio_token_id, cfg.audio_start_token_id, cfg.audio_end_token_id) == (151676, 151669, 151670)
    assert isinstance(cfg.audio_encoder, AudioEncoderConfig) and isinstance(cfg.text_decoder, TextDecoderConfig)
    assert cfg.audio_encoder.d_model == 1024 and cfg.audio_encoder.num_mel_bins == 128
    assert cfg.text_decoder.hidden_size == 2048 and cfg.text_decoder.vocab_size == 151936
    (tmp_path / "config.json").write_text(json.dumps(SAMPLE))
    loaded = ModelConfig.from_pretrained(tmp_path)
    assert loaded == cfg
    assert ModelConfig().audio_encoder.output_dim == ModelConfig().text_decoder.hidden_size == 2048  # the hot path's output contract
