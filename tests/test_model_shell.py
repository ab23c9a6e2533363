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
