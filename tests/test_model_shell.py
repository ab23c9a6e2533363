"""CPU: the Qwen3ASR shell's host logic (long-audio splitter pinned to the reference's own function;
reference tests/test_model.py:83-122)."""
import os

import numpy as np
import pytest

from qwen3_asr_mlx_b200.model import LANGUAGE_MAP, Qwen3ASR, TranscriptionResult, _find_split_points


def test_split_points_match_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "split_points_reference.npz"))
    names = sorted({k[0] for k in g.files})
    assert len(names) == 5
    for n in names:
        size, chunk, search, seed = (int(v) for v in g[n + "_args"])
        x = (np.random.default_rng(seed).standard_normal(size) * np.abs(np.sin(np.arange(size) / 9000.0))).astype(np.float32)
        assert _find_split_points(x, chunk, search) == list(g[n + "_points"]), n


def test_split_points_and_energies_match_reference_golden_gapped(golden_dir):
    """Config-4-style inputs (near-silent gaps), non-default frame sizes and chunk < search window."""
    g = np.load(os.path.join(golden_dir, "split_energy_reference.npz"))
    for name in sorted({k.split("_")[0] for k in g.files}):
        n, chunk, search, frame, seed = (int(v) for v in g[name + "_args"])
        r = np.random.default_rng(seed)
        x = (0.1 * r.standard_normal(n)).astype(np.float32)
        pos = 0
        while pos < n:
            pos += int(r.uniform(3.0, 6.0) * 16000)
            x[pos: pos + 8000] *= np.float32(1e-3)
        assert _find_split_points(x, chunk, search, frame) == list(g[name + "_points"]), name
        nf = n // frame
        if nf:
            e = np.sqrt(np.mean(x[: nf * frame].reshape(nf, frame) ** 2, axis=1)).astype(np.float32)
            assert np.array_equal(e.view(np.uint32), g[name + "_energy"].view(np.uint32)), name


def test_split_points_reference_known_answers():
    sr = 16_000
    assert _find_split_points(np.zeros(sr, dtype=np.float32), sr * 20, 5 * sr) == []
    x = np.random.default_rng(0).random(int(sr * 10 * 2.5)).astype(np.float32)
    pts = _find_split_points(x, sr * 10, 5 * sr)
    assert len(pts) == 2 and pts == sorted(pts) and all(0 <= p < len(x) for p in pts)
    x = np.ones(sr * 25, dtype=np.float32) * 0.5
    x[sr * 9: sr * 11] = 0.0
    pts = _find_split_points(x, sr * 10, sr * 5)
    assert len(pts) >= 1 and sr * 9 <= pts[0] <= sr * 11  # snaps into the silent region


def test_result_and_language_helpers():
    r = TranscriptionResult(text="hi", language="English", duration=1.0)
    assert (r.text, r.language, r.duration) == ("hi", "English", 1.0)
    shell = object.__new__(Qwen3ASR)
    assert shell._resolve_language(None) == "English" and shell._resolve_language("auto") == "English"
    assert shell._resolve_language("de") == LANGUAGE_MAP["de"] == "German"
    assert shell._resolve_language("Klingon") == "Klingon"
    with pytest.raises(ValueError):
        Qwen3ASR._as_samples(np.zeros((2, 100), dtype=np.float32))


def _reference_split_fn():
    """The reference's own _find_split_points, extracted with ast (its module imports mlx); authoring container only."""
    import ast

    path = "/root/reference/src/qwen3_asr_mlx/model.py"
    if not os.path.exists(path):
        pytest.skip("reference tree not present (GPU box)")
    node = next(n for n in ast.parse(open(path).read()).body if isinstance(n, ast.FunctionDef) and n.name == "_find_split_points")
    ns = {"np": np}
    exec(compile(ast.Module(body=[node], type_ignores=[]), "reference_model_py", "exec"), ns)
    return ns["_find_split_points"]


def test_split_points_equal_the_reference_function_on_random_inputs():
    """Beyond the committed golden cases: 150 random (length, chunk, search, frame) combinations, including chunk < search,
    frames that do not divide the length, silence and constant segments (ties -> first minimum)."""
    ref = _reference_split_fn()
    rng = np.random.default_rng(2026)
    for case in range(150):
        n = int(rng.integers(1, 200_000))
        frame = int(rng.choice([480, 480, 480, 160, 1000, 37]))
        chunk = int(rng.integers(max(1, n // 20), max(2, n)))
        search = int(rng.integers(0, 3 * chunk))
        x = (rng.standard_normal(n) * np.abs(np.sin(np.arange(n) / rng.uniform(500, 9000)))).astype(np.float32)
        if case % 5 == 0:
            a = int(rng.integers(0, n))
            x[a: a + int(rng.integers(1, 5000))] = 0.0          # exact ties inside a silent stretch
        if case % 7 == 0:
            x[:] = np.float32(0.25)                             # constant signal: every frame ties
        assert _find_split_points(x, chunk, search, frame) == ref(x, chunk, search, frame), (case, n, chunk, search, frame)


def test_language_map_covers_reference_hints():
    # hints the round-1 table lacked (ADVICE): they must resolve to the full names the reference prompts use
    shell = object.__new__(Qwen3ASR)
    for code, name in {"da": "Danish", "fi": "Finnish", "hu": "Hungarian", "bg": "Bulgarian", "ro": "Romanian", "ta": "Tamil",
                       "ur": "Urdu", "af": "Afrikaans", "DA": "Danish"}.items():
        assert shell._resolve_language(code) == name
    assert len(LANGUAGE_MAP) == 67


def test_hub_ids_resolve_through_huggingface_hub(tmp_path, monkeypatch):
    """A non-directory argument is a hub repo id: snapshot_download / hf_hub_download are imported lazily and called like
    the reference does (model.py:170-176, config.py:139-148, encoder.py:342-344)."""
    import json

    import huggingface_hub

    from qwen3_asr_mlx_b200 import _hub
    from qwen3_asr_mlx_b200.config import AudioEncoderConfig, ModelConfig, TextDecoderConfig

    (tmp_path / "config.json").write_text(json.dumps({"audio_encoder_config": {"d_model": 256, "encoder_layers": 3}, "text_config": {}}))
    calls = []
    monkeypatch.setattr(huggingface_hub, "snapshot_download", lambda repo_id, **kw: calls.append(("snap", repo_id, kw)) or str(tmp_path))
    monkeypatch.setattr(huggingface_hub, "hf_hub_download", lambda repo_id, filename, **kw: calls.append(("file", repo_id, filename)) or str(tmp_path / filename))
    assert _hub.model_dir(tmp_path) == tmp_path and not calls
    assert _hub.model_dir("org/some-model", revision="main") == tmp_path
    assert calls == [("snap", "org/some-model", {"revision": "main"})]
    cfg = AudioEncoderConfig.from_pretrained("org/some-model")
    assert cfg.d_model == 256 and cfg.encoder_layers == 3 and calls[-1] == ("file", "org/some-model", "config.json")
    assert ModelConfig.from_pretrained("org/some-model").audio_encoder.d_model == 256
    assert TextDecoderConfig.from_pretrained(tmp_path).hidden_size == TextDecoderConfig().hidden_size


def test_load_audio_falls_back_on_any_fast_path_failure(tmp_path, monkeypatch):
    """A WAV whose fmt chunk is truncated makes the native reader fail with struct.error; like the reference
    (audio.py:189-193: `except Exception`) the loader then tries soundfile instead of propagating."""
    import struct
    import sys
    import types

    from qwen3_asr_mlx_b200.audio import load_audio

    body = b"fmt " + struct.pack("<I", 8) + b"\x01\x00\x01\x00\x80\x3e\x00\x00" + b"data" + struct.pack("<I", 4) + b"\0\0\0\0"
    path = tmp_path / "bad.wav"
    path.write_bytes(b"RIFF" + struct.pack("<I", 4 + len(body)) + b"WAVE" + body)
    fake = types.ModuleType("soundfile")
    fake.read = lambda p, dtype, always_2d: (np.full((10, 2), 0.5, dtype=np.float32), 16000)
    monkeypatch.setitem(sys.modules, "soundfile", fake)
    out = load_audio(path)
    assert out.shape == (10,) and np.allclose(out, 0.5)
