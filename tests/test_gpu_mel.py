"""GPU parity of the mel frontend (through the C ABI) against the reference's golden vectors and
the oracle.  Tolerance: log-mel max-abs error <= 1e-4 (BASELINE.json north_star)."""
import ctypes
import os

import numpy as np
import pytest

from helpers import MEL_TOL, synth
from oracle import mel_np

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def audio_mod():
    from qwen3_asr_mlx_b200 import audio

    return audio


def test_matches_reference_golden_vectors(audio_mod, golden_dir):
    g = np.load(os.path.join(golden_dir, "mel_reference.npz"))
    names = [k[3:] for k in g.files if k.startswith("in_")]
    for name in names:
        got = np.array(audio_mod.log_mel_spectrogram(g["in_" + name]))
        ref = g["out_" + name]
        assert got.shape == ref.shape and got.dtype == np.float32
        assert np.abs(got - ref).max() <= MEL_TOL, name


@pytest.mark.parametrize("n", [160, 161, 319, 320, 1600, 16000 - 1, 16000, 51234, 160000, 480000])
def test_matches_oracle_on_seeded_audio(audio_mod, n):
    x = synth(np.random.default_rng(n), n)
    got = np.array(audio_mod.log_mel_spectrogram(x))
    assert got.shape == (128, n // 160)
    assert np.abs(got - mel_np.log_mel_spectrogram_fast(x)).max() <= MEL_TOL


def test_reference_known_answers(audio_mod):
    # reference tests/test_audio.py:49-56, 71-78, 80-95
    from qwen3_asr_mlx_b200 import DeviceArray

    mel = audio_mod.log_mel_spectrogram(np.zeros(16000, dtype=np.float32))
    assert isinstance(mel, DeviceArray) and mel.shape == (128, 100)
    assert np.allclose(np.array(mel), -1.5, atol=1e-3)
    t = np.linspace(0.0, 1.0, 16000, endpoint=False)
    tone = np.array(audio_mod.log_mel_spectrogram(np.sin(2 * np.pi * 440 * t).astype(np.float32)))
    assert tone.max() < 10.0 and tone.min() > -5.0
    for d in (0.5, 2.0, 5.0):
        assert audio_mod.log_mel_spectrogram(np.zeros(int(d * 16000), dtype=np.float32)).shape == (128, int(d * 100))


def test_error_behaviour(audio_mod):
    with pytest.raises(ValueError):  # the reference fails with numpy's zero-size reduction ValueError
        audio_mod.log_mel_spectrogram(np.zeros(159, dtype=np.float32))
    with pytest.raises(ValueError):  # model.py:298-301
        audio_mod.log_mel_spectrogram(np.zeros((2, 16000), dtype=np.float32))
    with pytest.raises(ValueError):
        audio_mod.log_mel_spectrogram(np.zeros(16000, dtype=np.float32), n_fft=512)
    with pytest.raises(ValueError):
        audio_mod.log_mel_spectrogram_batch([])


def test_ragged_batch_equals_loop_of_singles(audio_mod):
    rng = np.random.default_rng(11)
    xs = [synth(rng, int(n)) for n in [160, 16000, 47999, 300001, 1234, 160 * 33 + 7, 99999]]
    mel, foffs = audio_mod.log_mel_spectrogram_batch(xs)
    m = np.array(mel)
    assert list(np.diff(foffs)) == [len(x) // 160 for x in xs]
    for u, x in enumerate(xs):
        single = np.array(audio_mod.log_mel_spectrogram(x))
        blk = m[128 * int(foffs[u]): 128 * int(foffs[u + 1])].reshape(128, -1)
        assert np.array_equal(blk, single)  # same kernels, same per-utterance max: bit-identical
        assert np.abs(single - mel_np.log_mel_spectrogram_fast(x)).max() <= MEL_TOL


def test_host_pointer_entry_point(audio_mod):
    from qwen3_asr_mlx_b200 import runtime

    h = runtime.frontend_handle()
    xs = [synth(np.random.default_rng(3), 32000), synth(np.random.default_rng(4), 8000)]
    packed = np.concatenate(xs)
    soffs = np.array([0, 32000, 40000], dtype=np.int64)
    out = np.empty(128 * (200 + 50), dtype=np.float32)
    h.check(h.lib.qasr_mel_host(h.ptr, ctypes.c_void_p(packed.ctypes.data), runtime.i64_ptr(soffs), 2, ctypes.c_void_p(out.ctypes.data)))
    assert np.abs(out[: 128 * 200].reshape(128, 200) - mel_np.log_mel_spectrogram_fast(xs[0])).max() <= MEL_TOL
    assert np.abs(out[128 * 200:].reshape(128, 50) - mel_np.log_mel_spectrogram_fast(xs[1])).max() <= MEL_TOL


def test_config2_size_batch_properties(audio_mod):
    """BASELINE config 2 size (64 x 30 s): size-independent properties + spot parity."""
    rng = np.random.default_rng(1)
    base = [synth(rng, 480000) for _ in range(4)]
    xs = [base[i % 4] * np.float32(1.0 / (1 + i // 4)) for i in range(64)]
    mel, foffs = audio_mod.log_mel_spectrogram_batch(xs)
    m = np.array(mel).reshape(64, 128, 3000)
    assert np.isfinite(m).all()
    # clamp property: after (x+4)/4 the dynamic range per utterance is at most 8/4 = 2
    rng_per_utt = m.reshape(64, -1).max(axis=1) - m.reshape(64, -1).min(axis=1)
    assert (rng_per_utt <= 2.0 + 1e-6).all()
    for u in (0, 37, 63):
        assert np.abs(m[u] - mel_np.log_mel_spectrogram_fast(xs[u])).max() <= MEL_TOL
    # scaling audio by a scales power by a^2: log-mel shifts by 2*log10(a)/4 where nothing is clamped
    shift = m[0] - m[4]
    unclamped = (m[0] > m[0].max() - 1.5) & (m[4] > m[4].max() - 1.5)
    assert np.abs(shift[unclamped] - 2 * np.log10(2.0) / 4).max() <= 2e-4


def test_config4_twenty_minute_utterance(audio_mod):
    """BASELINE config 4: one 20-minute utterance, single pass, per-call global max."""
    rng = np.random.default_rng(4)
    x = (0.1 * rng.standard_normal(19_200_000)).astype(np.float32)
    x[3_000_000:3_008_000] *= 1e-3  # a near-silent gap
    got = np.array(audio_mod.log_mel_spectrogram(x))
    assert got.shape == (128, 120000)
    assert np.abs(got - mel_np.log_mel_spectrogram_fast(x)).max() <= MEL_TOL


def test_non_finite_samples_poison_the_utterance_like_the_reference(audio_mod):
    """A NaN / Inf sample makes the reference's log-mel NaN for the WHOLE utterance (np.maximum and .max() propagate NaN,
    audio.py:274-275; checked against the reference run verbatim) -- and only for that utterance of a batch."""
    rng = np.random.default_rng(0)
    clean = (0.1 * rng.standard_normal(16000)).astype(np.float32)
    for bad in (np.nan, np.inf, -np.inf):
        x = clean.copy()
        x[5000] = bad
        mel, foffs = audio_mod.log_mel_spectrogram_batch([clean, x, clean])
        m = np.array(mel).reshape(-1)
        blocks = [m[128 * int(foffs[u]): 128 * int(foffs[u + 1])] for u in range(3)]
        assert np.isnan(blocks[1]).all()
        assert np.isfinite(blocks[0]).all() and np.array_equal(blocks[0], blocks[2])


def test_batch_packing_paths_agree(audio_mod):
    """Varlen packing without a per-utterance host loop: a list of host arrays (one H2D), device tensors lying back to back
    in one allocation (zero copy), scattered / misaligned device tensors (qasr_pack_audio: one table upload + one kernel) and a
    host/device mix all give the bit-identical packed log-mel, and a 1024-utterance batch is a single pack launch."""
    import torch

    from qwen3_asr_mlx_b200 import runtime

    rng = np.random.default_rng(21)
    lens = [160, 16001, 47999, 300003, 1234, 160 * 33 + 7, 99998, 16000]
    xs = [synth(rng, n) for n in lens]
    want, foffs = audio_mod.log_mel_spectrogram_batch(xs)                       # host arrays
    want = np.array(want)
    scattered = [torch.from_numpy(x).cuda() for x in xs]                         # separate allocations
    flat = torch.from_numpy(np.concatenate(xs)).cuda()
    adjacent = list(flat.split(lens))                                            # views lying back to back
    odd = torch.zeros(sum(lens) + 3 * len(lens), device="cuda")
    misaligned, pos = [], 0
    for x in xs:                                                                 # every segment starts 4 bytes off a 16 B boundary
        pos += (-pos) % 4 + 1
        odd[pos: pos + len(x)] = torch.from_numpy(x).cuda()
        misaligned.append(odd[pos: pos + len(x)])
        pos += len(x)
    mixed = [scattered[i] if i % 2 else xs[i] for i in range(len(xs))]
    for name, batch in (("scattered", scattered), ("adjacent", adjacent), ("misaligned", misaligned), ("mixed", mixed)):
        got, fo = audio_mod.log_mel_spectrogram_batch(batch)
        assert np.array_equal(fo, foffs) and np.array_equal(np.array(got), want), name
    h = runtime.frontend_handle()
    many = [torch.from_numpy(synth(rng, 16000)).cuda() for _ in range(8)] * 128  # 1024 one-second utterances
    l0 = h.stats()["kernel_launches"]
    got, fo = audio_mod.log_mel_spectrogram_batch(many)
    assert h.stats()["kernel_launches"] - l0 == 3                                # pack + log-mel + normalise
    g = np.array(got).reshape(1024, 128, 100)
    assert np.array_equal(g[:8], g[8:16]) and np.array_equal(g[3], np.array(audio_mod.log_mel_spectrogram(many[3])))
