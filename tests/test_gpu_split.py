"""GPU parity of the long-audio splitter (reference _find_split_points, model.py:454-513) through the C ABI:
frame energies bit-identical to the reference's numpy expression, cut positions identical to the reference's own
function (tests/golden/split_*_reference.npz were produced by executing it, oracle/gen_golden.py)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def enc():
    from qwen3_asr_mlx_b200 import AudioEncoder, AudioEncoderConfig

    e = AudioEncoder(AudioEncoderConfig(d_model=256, encoder_layers=1, encoder_attention_heads=4, encoder_ffn_dim=512, output_dim=256))
    yield e
    e.close()


def _gapped(seed, n):
    r = np.random.default_rng(seed)
    x = (0.1 * r.standard_normal(n)).astype(np.float32)
    pos = 0
    while pos < n:
        pos += int(r.uniform(3.0, 6.0) * 16000)
        x[pos: pos + 8000] *= np.float32(1e-3)
    return x


def test_energy_and_points_match_reference_golden(enc, golden_dir):
    import torch

    g = np.load(os.path.join(golden_dir, "split_energy_reference.npz"))
    names = sorted({k.split("_")[0] for k in g.files})
    assert names == ["p", "q", "r", "s", "t"]
    for name in names:
        n, chunk, search, frame, seed = (int(v) for v in g[name + "_args"])
        x = torch.from_numpy(_gapped(seed, n)).cuda()
        pts, energy = enc.find_split_points(x, chunk, search, frame, return_energy=True)
        assert np.array_equal(energy.cpu().numpy().view(np.uint32), g[name + "_energy"].view(np.uint32)), name  # bit-exact
        assert pts == list(g[name + "_points"]), name


def test_points_match_reference_golden_noise_cases(enc, golden_dir):
    import torch

    g = np.load(os.path.join(golden_dir, "split_points_reference.npz"))
    for n in sorted({k[0] for k in g.files}):
        size, chunk, search, seed = (int(v) for v in g[n + "_args"])
        x = (np.random.default_rng(seed).standard_normal(size) * np.abs(np.sin(np.arange(size) / 9000.0))).astype(np.float32)
        assert enc.find_split_points(torch.from_numpy(x).cuda(), chunk, search) == list(g[n + "_points"]), n


def test_edge_cases(enc):
    import torch

    sr = 16000
    assert enc.find_split_points(torch.zeros(sr, device="cuda"), sr * 20, 5 * sr) == []          # shorter than one chunk
    assert enc.find_split_points(torch.zeros(100, device="cuda"), 10, 10) == []                   # no whole frame (model.py:486)
    assert enc.find_split_points(torch.zeros(0, device="cuda"), 10, 10) == []
    z = torch.zeros(sr * 25, device="cuda")                                                       # digital silence: ties -> first frame of the window
    assert enc.find_split_points(z, sr * 10, sr * 5) == [((sr * 10) // 480 - (sr * 5) // 480) * 480, ((sr * 20) // 480 - (sr * 5) // 480) * 480]
    x = torch.ones(sr * 25, device="cuda") * 0.5
    x[sr * 9: sr * 11] = 0.0
    pts = enc.find_split_points(x, sr * 10, sr * 5)
    assert sr * 9 <= pts[0] <= sr * 11                                                            # reference tests/test_model.py:104-122
    assert enc.find_split_points(x, sr * 10, 0) == [sr * 10, sr * 20]                             # degenerate window -> the boundary itself
    with pytest.raises(ValueError):
        enc.find_split_points(x, 0, 10)
    with pytest.raises(ValueError):
        enc.find_split_points(x.cpu(), 10, 10)


def test_twenty_minute_file_equals_host_twin(enc):
    """Config 4 size (19.2 M samples): device cuts == the numpy twin (itself pinned to the reference's function)."""
    import torch

    from qwen3_asr_mlx_b200.model import _find_split_points

    x = _gapped(4, 1200 * 16000)
    pts = enc.find_split_points(torch.from_numpy(x).cuda(), 30 * 16000, 5 * 16000)
    assert len(pts) == 39 and pts == _find_split_points(x, 30 * 16000, 5 * 16000)


def test_transcribe_long_audio_segments(enc):
    """transcribe() with chunk_duration < duration: device splitter + one varlen batch; the decoder backend sees every segment."""
    from qwen3_asr_mlx_b200 import Qwen3ASR
    from qwen3_asr_mlx_b200.model import _find_split_points

    seen = []

    def backend(emb, n_tokens, lang, max_tokens, **kw):
        seen.append((int(n_tokens), tuple(emb.shape), max_tokens))
        return f"seg{len(seen)}"

    shell = Qwen3ASR(enc.config, enc, decoder_backend=backend)
    x = _gapped(9, 16000 * 95 + 123)
    res = shell.transcribe(x, chunk_duration=30.0)
    cuts = _find_split_points(x, 30 * 16000, 5 * 16000)
    bounds = [0] + cuts + [len(x)]
    assert len(seen) == len(bounds) - 1 == 4
    for (ntok, shape, mt), a, b in zip(seen, bounds[:-1], bounds[1:]):
        assert ntok == enc.num_tokens((b - a) // 160) and shape == (ntok, enc.config.output_dim)
        assert mt == max(256, int((b - a) / 16000 * 50))  # model.py:415-416
    assert res.text == "seg1 seg2 seg3 seg4" and abs(res.duration - len(x) / 16000) < 1e-9
    # per-segment mel max (model.py:418): a segment encoded alone gives the same embeddings
    emb, toffs, spans = shell.encode_long(x, 30.0)
    alone = np.array(enc.encode_audio_batch([x[spans[2][0]: spans[2][1]]])[0])
    assert np.array_equal(np.array(emb)[int(toffs[2]): int(toffs[3])], alone)


def test_random_inputs_equal_host_twin(enc):
    """80 random (length, chunk, search, frame) combinations, with silent stretches and constant signals (exact ties), against
    the numpy twin -- which tests/test_model_shell.py pins to the reference's own function on the same kind of inputs."""
    import torch

    from qwen3_asr_mlx_b200.model import _find_split_points

    rng = np.random.default_rng(2027)
    for case in range(80):
        n = int(rng.integers(1, 300_000))
        frame = int(rng.choice([480, 480, 480, 160, 1000, 37, 7, 129]))
        chunk = int(rng.integers(max(1, n // 20), max(2, n)))
        search = int(rng.integers(0, 3 * chunk))
        x = (rng.standard_normal(n) * np.abs(np.sin(np.arange(n) / rng.uniform(500, 9000)))).astype(np.float32)
        if case % 5 == 0:
            a = int(rng.integers(0, n))
            x[a: a + int(rng.integers(1, 5000))] = 0.0
        if case % 7 == 0:
            x[:] = np.float32(0.25)
        pts, energy = enc.find_split_points(torch.from_numpy(x).cuda(), chunk, search, frame, return_energy=True)
        nf = n // frame
        if nf:
            want_e = np.sqrt(np.mean(x[: nf * frame].reshape(nf, frame) ** 2, axis=1)).astype(np.float32)
            assert np.array_equal(energy.cpu().numpy().view(np.uint32), want_e.view(np.uint32)), (case, n, frame)
        assert pts == _find_split_points(x, chunk, search, frame), (case, n, chunk, search, frame)
