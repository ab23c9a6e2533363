"""CPU: host-side logic and the C-ABI surface (no compute calls; there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

from qwen3_asr_mlx_b200 import _lib, launcher, weights
from qwen3_asr_mlx_b200.config import AudioEncoderConfig
from qwen3_asr_mlx_b200.encoder import AudioEncoder

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------ C ABI
def header_symbols():
    found = set()
    for header in ("qasr.h", "qasr_decoder.h"):  # every header under include/
        text = open(os.path.join(ROOT, "include", header)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        found |= set(re.findall(r"\b(qasr_[a-z0-9_]+)\s*\(", text))
    return sorted(found)


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"libqasr.so does not export {name}"
    assert sorted(_lib.exported_symbols()) == declared, "ctypes signature table out of sync with include/*.h"


def test_default_config_matches_reference_defaults():
    lib = _lib.load()
    c = _lib.QasrConfig()
    lib.qasr_default_config(ctypes.byref(c))
    d = AudioEncoderConfig()
    for f, _ in _lib.QasrConfig._fields_:
        assert getattr(c, f) == getattr(d, f), f


def test_count_frames_and_tokens_rules():
    lib = _lib.load()
    out = ctypes.c_int64()
    assert lib.qasr_count_frames(159, ctypes.byref(out)) == _lib.QASR_ERR_INVALID
    assert "160" in _lib.last_error()
    for n in (160, 16000, 480000, 19_200_000, 12345):
        assert lib.qasr_count_frames(n, ctypes.byref(out)) == 0 and out.value == n // 160
    for T, tok in ((100, 13), (300, 39), (250, 33), (50, 7), (1, 1), (1000, 130), (3000, 390), (120000, 15600)):
        assert lib.qasr_count_tokens(None, T, ctypes.byref(out)) == 0 and out.value == tok
        assert launcher.tokens_for_samples(T * 160) == tok


def test_constant_tables_match_oracle(golden_dir):
    lib = _lib.load()
    fb = np.empty((128, 201), dtype=np.float32)
    assert lib.qasr_mel_filterbank(fb.ctypes.data_as(ctypes.POINTER(ctypes.c_float))) == 0
    assert np.array_equal(fb, np.load(os.path.join(golden_dir, "mel_filterbank.npy")))
    win = np.empty(400, dtype=np.float32)
    assert lib.qasr_hann_window(win.ctypes.data_as(ctypes.POINTER(ctypes.c_float))) == 0
    assert np.array_equal(win, np.hanning(400).astype(np.float32))  # symmetric Hann, audio.py:222


def test_no_cpu_fallback():
    """Without a GPU the product refuses to run instead of computing on the host."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    c = _lib.QasrConfig()
    lib.qasr_default_config(ctypes.byref(c))
    h = ctypes.c_void_p()
    assert lib.qasr_create(0, ctypes.byref(c), ctypes.byref(h)) == _lib.QASR_ERR_UNSUPPORTED
    assert "no CPU fallback" in _lib.last_error()
    from qwen3_asr_mlx_b200 import log_mel_spectrogram

    with pytest.raises(_lib.QasrError):
        log_mel_spectrogram(np.zeros(16000, dtype=np.float32))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "qwen3_asr_mlx_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"


def test_status_to_exception_mapping():
    with pytest.raises(ValueError):
        _lib.check(_lib.QASR_ERR_INVALID)
    with pytest.raises(MemoryError):
        _lib.check(_lib.QASR_ERR_NOMEM)
    with pytest.raises(_lib.QasrError):
        _lib.check(_lib.QASR_ERR_CUDA)


# ------------------------------------------------------------------ config / weights
def test_config_from_dict_rules():
    # reference config.py:31-58
    c = AudioEncoderConfig.from_dict({"audio_encoder_config": {"d_model": 512, "num_hidden_layers": 6}})
    assert c.d_model == 512 and c.encoder_layers == 6 and c.encoder_ffn_dim == 4096
    c = AudioEncoderConfig.from_dict({"encoder_layers": 3, "num_hidden_layers": 9, "output_dim": 1024})
    assert c.encoder_layers == 3 and c.output_dim == 1024
    assert AudioEncoderConfig.from_dict({}) == AudioEncoderConfig()


def test_parameter_inventory_of_1p7b():
    shapes = dict(weights.parameter_shapes(AudioEncoderConfig()))
    assert sum(int(np.prod(s)) for s in shapes.values()) == 317_477_504
    assert shapes["conv2d2.weight"] == (480, 3, 3, 480) and shapes["conv_out.weight"] == (1024, 7680)
    assert "conv_out.bias" not in shapes  # encoder.py:174-178: bias=False
    per_layer = sum(int(np.prod(s)) for k, s in shapes.items() if k.startswith("layers.0."))
    assert per_layer == 12_596_224


def test_random_init_is_seeded_and_bounded(tmp_path):
    cfg = AudioEncoderConfig(d_model=128, encoder_layers=1, encoder_attention_heads=2, encoder_ffn_dim=256, output_dim=64)
    a, b, c = weights.random_init(cfg, 3), weights.random_init(cfg, 3), weights.random_init(cfg, 4)
    assert all(np.array_equal(a[k], b[k]) for k in a) and not np.array_equal(a["proj1.weight"], c["proj1.weight"])
    assert np.abs(a["layers.0.fc2.weight"]).max() <= 1 / np.sqrt(256) and np.abs(a["layers.0.fc2.bias"]).max() <= 1 / np.sqrt(256)
    assert np.abs(a["conv2d2.weight"]).max() <= 1 / np.sqrt(480 * 9) and not a["conv2d2.bias"].any()
    assert (a["ln_post.weight"] == 1).all() and not a["ln_post.bias"].any()
    # safetensors round trip with the reference's "audio_tower." prefix (encoder.py:349-356)
    path = tmp_path / "model.safetensors"
    extra = dict(a)
    weights.save_safetensors(extra, path)
    from safetensors.numpy import load_file

    raw = load_file(str(path))
    assert all(k.startswith("audio_tower.") for k in raw)
    back = weights.load_safetensors(path)
    assert set(back) == set(a) and all(np.array_equal(back[k], a[k]) for k in a)


# ------------------------------------------------------------------ encoder host helpers (reference tests/test_encoder.py:126-164)
def test_block_attention_mask_helper():
    assert AudioEncoder._block_attention_mask(13, [0, 13]) is None
    m = AudioEncoder._block_attention_mask(26, [0, 13, 26])
    assert m.shape == (1, 1, 26, 26)
    assert (m[0, 0, :13, :13] == 0).all() and (m[0, 0, 13:, 13:] == 0).all()
    assert (m[0, 0, :13, 13:] < -1e8).all() and (m[0, 0, 13:, :13] < -1e8).all()


def test_conv_output_length_helper():
    assert AudioEncoder._conv_output_length(100) == 13 and AudioEncoder._conv_output_length(50) == 7


# ------------------------------------------------------------------ launcher
def test_lpt_partition_balances_config3():
    # BASELINE.json config 3: 4096 utterances of U[1,30] s, length seed 20261018
    lengths = np.random.default_rng(20261018).integers(16000, 480001, size=4096)
    assert int(lengths.sum()) == 1_012_290_886
    costs = [launcher.tokens_for_samples(int(n)) for n in lengths]
    assert sum(costs) == 823_054
    parts = launcher.lpt_partition(costs, 8)
    assert sorted(i for p in parts for i in p) == list(range(4096))
    loads = [sum(costs[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= 13 * 30  # within one utterance
    assert parts == launcher.lpt_partition(costs, 8)  # deterministic: every rank derives the same plan


def test_split_by_budget():
    costs = [390] * 10
    subs = launcher.split_by_budget(list(range(10)), costs, 1000)
    assert subs == [[0, 1], [2, 3], [4, 5], [6, 7], [8, 9]]
    assert launcher.split_by_budget([3], costs, 10) == [[3]]  # a single over-budget utterance still runs


def test_encode_sharded_single_rank_restores_order():
    import torch

    n_samples = [16000 * k for k in (3, 1, 7, 2, 5)]

    def fake_encode(idx):
        rows = [torch.full((launcher.tokens_for_samples(n_samples[i]), 4), float(i)) for i in idx]
        offs = np.cumsum([0] + [r.shape[0] for r in rows])
        return torch.cat(rows), offs

    emb, offs, mine = launcher.encode_sharded(fake_encode, n_samples, 4, rank=0, world_size=1, tokens_per_call=60)
    assert mine == [0, 1, 2, 3, 4]
    for i in range(5):
        assert (emb[int(offs[i]): int(offs[i + 1])] == float(i)).all()


def test_load_audio_matches_reference_golden(tmp_path, golden_dir):
    """WAV decode + linear-interpolation resample == the reference's load_audio (audio.py:103-204), executed verbatim by
    oracle/gen_golden.py on the same bytes."""
    import os

    from helpers import WAV_CASES, make_wav
    from qwen3_asr_mlx_b200.audio import load_audio

    g = np.load(os.path.join(golden_dir, "load_audio_reference.npz"))
    assert sorted(g.files) == sorted(WAV_CASES)
    for name, kw in WAV_CASES.items():
        path = tmp_path / (name + ".wav")
        path.write_bytes(make_wav(**kw))
        got = load_audio(path)
        assert got.dtype == np.float32 and got.ndim == 1
        assert np.array_equal(got, g[name]), name
