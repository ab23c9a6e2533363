"""GPU parity of the audio encoder (through the C ABI) against the oracle.
Tolerance: embedding relative (Frobenius) error <= 2e-2, bf16 kernels vs the fp32 oracle
(BASELINE.json north_star).  Shape contract mirrored from reference tests/test_encoder.py."""
import ctypes
import os

import numpy as np
import pytest

from helpers import EMB_TOL, bf16_bits, bf16_round, rel_err, synth
from oracle import encoder_torch, mel_np

pytestmark = pytest.mark.gpu


def _small():
    from qwen3_asr_mlx_b200 import AudioEncoderConfig

    return AudioEncoderConfig(d_model=256, encoder_layers=2, encoder_attention_heads=4, encoder_ffn_dim=512, output_dim=256)


@pytest.fixture(scope="module")
def small():
    from qwen3_asr_mlx_b200 import AudioEncoder, weights

    cfg = _small()
    params = weights.random_init(cfg, seed=7, exercise_all=True)
    enc = AudioEncoder(cfg)
    enc.load_weights(params)
    yield cfg, params, enc
    enc.close()


@pytest.fixture(scope="module")
def full():
    from qwen3_asr_mlx_b200 import AudioEncoder, AudioEncoderConfig, weights

    cfg = AudioEncoderConfig()
    params = weights.random_init(cfg, seed=1234)
    enc = AudioEncoder(cfg)
    enc.load_weights(params)
    yield cfg, params, enc
    enc.close()


@pytest.mark.parametrize("pair", [0, 16], ids=["cta_group1", "cta_group2"])
@pytest.mark.parametrize("shape", [(128, 256, 64), (1, 256, 64), (300, 1024, 1024), (130, 2048, 4096), (1000, 3072, 1024)])
def test_tcgen05_gemm_kernel(shape, pair):
    from qwen3_asr_mlx_b200 import _lib

    M, N, K = shape
    rng = np.random.default_rng(M + N + K)
    a = bf16_round(rng.standard_normal((M, K)).astype(np.float32))
    w = bf16_round((rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32))
    bias = rng.standard_normal(N).astype(np.float32)
    out = np.zeros((M, N), dtype=np.float32)
    lib = _lib.load()
    u16, f32 = ctypes.POINTER(ctypes.c_uint16), ctypes.POINTER(ctypes.c_float)
    ab, wb = bf16_bits(a), bf16_bits(w)
    _lib.check(lib.qasr_test_gemm(0, ab.ctypes.data_as(u16), wb.ctypes.data_as(u16), bias.ctypes.data_as(f32), M, N, K, 0 + pair, out.ctypes.data_as(f32)))
    ref = a.astype(np.float64) @ w.astype(np.float64).T + bias
    assert np.abs(out - ref).max() <= 1e-4  # fp32 accumulation of exact bf16 products
    # residual epilogue (TMA reduce-add into an fp32 matrix that already holds values)
    resid = rng.standard_normal((M, N)).astype(np.float32)
    acc = resid.copy()
    _lib.check(lib.qasr_test_gemm(0, ab.ctypes.data_as(u16), wb.ctypes.data_as(u16), bias.ctypes.data_as(f32), M, N, K, 2 + pair, acc.ctypes.data_as(f32)))
    assert np.abs(acc - (ref + resid)).max() <= 1e-4


@pytest.mark.parametrize("n_samples", [160, 16000, 16000 * 4 + 7000, 16000 * 9 + 4321, 16000 * 20 + 333])
def test_small_config_parity_with_intermediates(small, n_samples):
    cfg, params, enc = small
    mel = mel_np.log_mel_spectrogram_fast(synth(np.random.default_rng(n_samples), n_samples))
    ref, inter = encoder_torch.encoder_forward(params, cfg, mel, return_intermediates=True)
    enc.set_debug(True)
    out = np.array(enc(mel))
    assert out.shape == (1,) + ref.shape
    for k in ("stem", "layer0", "hidden"):
        assert rel_err(enc.debug_read(k, ref.shape[0]), inter[k]) <= EMB_TOL, k
    enc.set_debug(False)
    assert rel_err(out[0], ref) <= EMB_TOL


def test_golden_anchor(small, golden_dir):
    cfg, params, enc = small
    g = np.load(os.path.join(golden_dir, "encoder_small.npz"))
    from qwen3_asr_mlx_b200 import log_mel_spectrogram

    out = np.array(enc(log_mel_spectrogram(g["audio"])))[0]  # CUDA mel feeding the CUDA encoder
    assert out.shape == g["emb"].shape
    assert rel_err(out, g["emb"]) <= EMB_TOL


@pytest.mark.parametrize("T,tokens", [(100, 13), (300, 39), (250, 33), (50, 7), (1, 1)])
def test_reference_shape_contract(small, T, tokens):
    # reference tests/test_encoder.py:64-97
    cfg, params, enc = small
    mel = np.random.default_rng(T).standard_normal((128, T)).astype(np.float32)
    out = enc(mel)
    assert out.shape == (1, tokens, cfg.output_dim)
    assert np.isfinite(np.array(out)).all()
    batched = enc(mel[None])  # (1, 128, T) accepted
    assert np.array_equal(np.array(batched), np.array(out))


def test_bad_input_shapes(small):
    cfg, params, enc = small
    with pytest.raises(ValueError):
        enc(np.zeros((64, 100), dtype=np.float32))
    with pytest.raises(ValueError):
        enc.encode_batch([])


def test_varlen_batch_equals_loop_of_singles(small):
    cfg, params, enc = small
    rng = np.random.default_rng(5)
    mels = [mel_np.log_mel_spectrogram_fast(synth(rng, int(n))) for n in (16000, 200000, 160 * 57, 16000 * 9, 160, 480000)]
    emb, toffs = enc.encode_batch(mels)
    e = np.array(emb)
    assert list(np.diff(toffs)) == [enc.num_tokens(m.shape[1]) for m in mels]
    for u, m in enumerate(mels):
        single = np.array(enc(m))[0]
        assert np.array_equal(e[int(toffs[u]): int(toffs[u + 1])], single)  # batching must not change results
        assert rel_err(single, encoder_torch.encoder_forward(params, cfg, m)) <= EMB_TOL


def test_sub_batching_is_transparent(small):
    """encode_audio_batch splits oversize batches into workspace-bounded sub-batches; results are unchanged."""
    cfg, params, enc = small
    rng = np.random.default_rng(17)
    xs = [synth(rng, int(n)) for n in rng.integers(4000, 120000, size=11)]
    whole, offs = enc.encode_audio_batch(xs)
    split, offs2 = enc.encode_audio_batch(xs, max_tokens_per_call=150)
    assert list(offs) == list(offs2)
    assert np.array_equal(np.array(whole), np.array(split))


def test_windows_are_independent(small):
    """Block-diagonal attention without a mask tensor: tokens of the first 8-s window do not depend on later audio."""
    cfg, params, enc = small
    mel = np.random.default_rng(3).standard_normal((128, 1000)).astype(np.float32)
    mel2 = mel.copy()
    mel2[:, 800:] += 1.0
    a, b = np.array(enc(mel))[0], np.array(enc(mel2))[0]
    assert a.shape == (130, cfg.output_dim)
    assert np.array_equal(a[:104], b[:104]) and not np.allclose(a[104:], b[104:])


def test_stem_grouping_is_transparent(small, monkeypatch):
    """The conv stem runs in bounded groups of chunks; a tiny group size must give identical results."""
    from qwen3_asr_mlx_b200 import AudioEncoder

    cfg, params, enc = small
    mel = mel_np.log_mel_spectrogram_fast(synth(np.random.default_rng(9), 16000 * 12 + 5000))
    ref = np.array(enc(mel))
    monkeypatch.setenv("QASR_STEM_GROUP", "3")
    enc2 = AudioEncoder(cfg)
    enc2.load_weights(params)
    assert np.array_equal(np.array(enc2(mel)), ref)
    enc2.close()


@pytest.mark.parametrize("env", [{"QASR_CTA_PAIR": "0"}, {"QASR_ATTN_TC": "0"}, {"QASR_CONV1_FP32": "1"},
                                 {"QASR_CTA_PAIR": "0", "QASR_ATTN_TC": "0", "QASR_GRAPHS": "0"},
                                 {"QASR_LANES": "2", "QASR_LANE_MIN_CHUNKS": "1"}, {"QASR_CONV_TAIL_SKIP": "0"}],
                         ids=["single_cta_gemm", "mma_sync_attention", "conv1_cuda_cores_fp32_weights", "all_alternative_kernels_eager",
                              "two_lanes_two_streams", "conv_tail_zero_mmas_issued"])
def test_kernel_variants_agree(small, monkeypatch, env):
    """The alternative kernels (cta_group::1 GEMM, mma.sync attention, eager launches) give the same
    embeddings as the default configuration (cta_group::2 GEMM, tcgen05 attention, graph replay)."""
    from qwen3_asr_mlx_b200 import AudioEncoder

    cfg, params, enc = small
    rng = np.random.default_rng(21)
    mels = [mel_np.log_mel_spectrogram_fast(synth(rng, int(n))) for n in (16000 * 11 + 999, 30000, 16000 * 30)]
    ref, _ = enc.encode_batch(mels)
    ref = np.array(ref)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    alt = AudioEncoder(cfg)
    alt.load_weights(params)
    for _ in range(3):  # eager, captured, replayed
        got = np.array(alt.encode_batch(mels)[0])
        assert rel_err(got, ref) <= 5e-3  # variants differ by bf16 rounding order (conv1 fp32 vs bf16 weights is the largest)
    for u, m in enumerate(mels):
        assert rel_err(np.array(alt(m))[0], encoder_torch.encoder_forward(params, cfg, m)) <= EMB_TOL
    alt.close()


@pytest.mark.parametrize("env", [{"QASR_PDL": "0"}, {"QASR_SERPENTINE": "0"}, {"QASR_L2_HINTS": "15"},
                                 {"QASR_PDL": "0", "QASR_SERPENTINE": "0", "QASR_GRAPHS": "0"}],
                         ids=["no_programmatic_dependent_launch", "rows_upwards_only", "l2_eviction_hints", "plain_stream_order_eager"])
def test_scheduling_knobs_change_no_bit(small, monkeypatch, env):
    """Programmatic dependent launch, the serpentine row order and the L2 eviction hints only change WHEN and in which order
    independent rows are computed: embeddings are bit-identical with every one of them switched."""
    from qwen3_asr_mlx_b200 import AudioEncoder

    cfg, params, enc = small
    rng = np.random.default_rng(22)
    audios = [synth(rng, int(n)) for n in (16000 * 13 + 5, 2400, 16000 * 30, 16000 * 4)]
    ref = np.array(enc.encode_audio_batch(audios)[0])
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    alt = AudioEncoder(cfg)  # qasr_create re-reads the switches
    alt.load_weights(params)
    for _ in range(3):  # eager, captured, replayed
        got = np.array(alt.encode_audio_batch(audios)[0])
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    alt.close()
    for k in env:
        monkeypatch.delenv(k)
    AudioEncoder(cfg).close()  # restores the process-wide PDL switch for the tests that follow


def test_graph_replay_matches_eager(small):
    """Same buffers three times: eager, graph capture, graph replay -> identical embeddings."""
    import torch

    cfg, params, enc = small
    x = synth(np.random.default_rng(33), 16000 * 17 + 77)
    audio = torch.from_numpy(x).cuda()
    soffs = np.array([0, len(x)], dtype=np.int64)
    out = torch.empty((enc.num_tokens(len(x) // 160), cfg.output_dim), dtype=torch.float32, device="cuda")
    results = []
    for _ in range(4):
        out.zero_()
        enc.encode_packed_audio(audio, soffs, out=out)
        results.append(out.cpu().numpy().copy())
    assert all(np.array_equal(results[0], r) for r in results[1:])
    audio.copy_(torch.from_numpy(synth(np.random.default_rng(34), len(x))).cuda())  # new content, same buffers: replayed graph must see it
    enc.encode_packed_audio(audio, soffs, out=out)
    assert not np.array_equal(out.cpu().numpy(), results[0])
    fresh = np.array(enc.encode_audio_batch([audio.cpu().numpy()])[0])
    assert np.array_equal(out.cpu().numpy(), fresh)


def test_full_arch_config1_parity(full):
    """BASELINE config 1: one 10 s utterance, Qwen3-ASR-1.7B encoder architecture, random init (seed 1234)."""
    cfg, params, enc = full
    x = synth(np.random.default_rng(0), 160000)
    mel = mel_np.log_mel_spectrogram_fast(x)
    ref = encoder_torch.encoder_forward(params, cfg, mel)
    out = np.array(enc(mel))
    assert out.shape == (1, 130, 2048)
    assert rel_err(out[0], ref) <= EMB_TOL
    emb, toffs = enc.encode_audio_batch([x])  # mel + encoder back to back on the device
    assert list(toffs) == [0, 130] and rel_err(np.array(emb), ref) <= EMB_TOL
    bf = np.array(enc.encode_audio_batch([x], out_dtype="bfloat16")[0])
    assert rel_err(bf, ref) <= EMB_TOL
    # per-utterance worst case over a few lengths (mixed-length, config 3 style)
    rng = np.random.default_rng(3)
    xs = [synth(rng, int(n)) for n in (16000, 77777, 250000)]
    emb, toffs = enc.encode_audio_batch(xs)
    e = np.array(emb)
    for u, xu in enumerate(xs):
        r = encoder_torch.encoder_forward(params, cfg, mel_np.log_mel_spectrogram_fast(xu))
        assert rel_err(e[int(toffs[u]): int(toffs[u + 1])], r) <= EMB_TOL


def test_full_arch_parity_with_upstream_implementation(full):
    """CUDA path vs the model authors' PyTorch audio tower (transformers Qwen3OmniMoeAudioEncoder, fp32 on the
    host), without going through oracle/: 12.5 s and 30 s of audio, full 24-layer 1.7B architecture."""
    upstream_hf = pytest.importorskip("upstream_hf")
    pytest.importorskip("transformers.models.qwen3_omni_moe.modeling_qwen3_omni_moe")
    cfg, params, enc = full
    model = upstream_hf.build_upstream(cfg, params)
    rng = np.random.default_rng(21)
    xs = [synth(rng, n) for n in (200000, 480000)]
    emb, toffs = enc.encode_audio_batch(xs)
    e = np.array(emb)
    for u, x in enumerate(xs):
        up = upstream_hf.upstream_forward(model, mel_np.log_mel_spectrogram_fast(x))
        got = e[int(toffs[u]): int(toffs[u + 1])]
        assert got.shape == up.shape
        assert rel_err(got, up) <= EMB_TOL


def test_full_arch_host_entry_point(full):
    cfg, params, enc = full
    rng = np.random.default_rng(8)
    xs = [synth(rng, 48000), synth(rng, 20000)]
    packed = np.concatenate(xs)
    soffs = np.array([0, 48000, 68000], dtype=np.int64)
    n_tok = enc.num_tokens(300) + enc.num_tokens(125)
    out = np.empty((n_tok, 2048), dtype=np.float32)
    toffs = enc.encode_audio_host(packed, soffs, out)
    dev, toffs2 = enc.encode_audio_batch(xs)
    assert list(toffs) == list(toffs2) and np.array_equal(out, np.array(dev))


def test_async_host_pipeline_matches_synchronous_call(small):
    """qasr_encode_audio_host_async / qasr_host_wait with both slots in flight give the synchronous results."""
    import torch

    cfg, params, enc = small
    rng = np.random.default_rng(12)
    batches = []
    for k in range(5):
        xs = [synth(rng, int(n)) for n in rng.integers(8000, 90000, size=3)]
        soffs = np.concatenate([[0], np.cumsum([len(x) for x in xs])]).astype(np.int64)
        batches.append((np.concatenate(xs), soffs))
    refs = []
    for audio, soffs in batches:
        n_tok = sum(enc.num_tokens(int(soffs[u + 1] - soffs[u]) // 160) for u in range(3))
        out = np.empty((n_tok, cfg.output_dim), dtype=np.float32)
        enc.encode_audio_host(audio, soffs, out)
        refs.append(out)
    pinned_in = [torch.empty(300000, dtype=torch.float32).pin_memory().numpy() for _ in range(2)]
    outs = [np.empty_like(r) for r in refs]
    for i, (audio, soffs) in enumerate(batches):
        s = i & 1
        enc.host_wait(s)
        pinned_in[s][: len(audio)] = audio
        enc.encode_audio_host_async(s, pinned_in[s][: len(audio)], soffs, outs[i])
    enc.host_wait(0)
    enc.host_wait(1)
    for got, ref in zip(outs, refs):
        assert np.array_equal(got, ref)
    with pytest.raises(ValueError):
        enc.encode_audio_host_async(2, batches[0][0], batches[0][1], outs[0])


def test_full_arch_config2_size_properties(full):
    """BASELINE config 2 size (64 x 30 s on one B200): shape, finiteness, batch == single, spot parity."""
    cfg, params, enc = full
    rng = np.random.default_rng(1)
    base = [synth(rng, 480000) for _ in range(4)]
    xs = [base[i % 4] for i in range(64)]
    emb, toffs = enc.encode_audio_batch(xs)
    e = np.array(emb)
    assert e.shape == (24960, 2048) and np.isfinite(e).all()
    assert list(np.diff(toffs)) == [390] * 64
    for u in range(4, 64):  # identical audio -> identical embeddings wherever it sits in the batch
        assert np.array_equal(e[390 * u: 390 * (u + 1)], e[390 * (u % 4): 390 * (u % 4 + 1)])
    single = np.array(enc.encode_audio_batch([xs[2]])[0])
    assert np.array_equal(single, e[780:1170])
    ref = encoder_torch.encoder_forward(params, cfg, mel_np.log_mel_spectrogram_fast(xs[2]))
    assert rel_err(single, ref) <= EMB_TOL


def test_from_pretrained_local_checkpoint_and_shell(tmp_path, small):
    """Qwen3ASR.from_pretrained on a local directory holding config.json + model.safetensors (audio_tower.* keys,
    bf16, MLX layouts) -- reference model.py:151-188, encoder.py:330-359 -- and the transcribe() shell behaviour."""
    import json

    import torch
    from safetensors.torch import save_file

    from qwen3_asr_mlx_b200 import AudioEncoder, Qwen3ASR, TranscriptionResult

    cfg, params, enc = small
    tensors = {"audio_tower." + k: torch.from_numpy(v).to(torch.bfloat16) for k, v in params.items()}
    tensors["model.embed_tokens.weight"] = torch.zeros(4, 4, dtype=torch.bfloat16)  # decoder tensors are ignored
    save_file(tensors, str(tmp_path / "model.safetensors"))
    (tmp_path / "config.json").write_text(json.dumps({"audio_encoder_config": {
        "d_model": cfg.d_model, "encoder_layers": cfg.encoder_layers, "encoder_attention_heads": cfg.encoder_attention_heads,
        "encoder_ffn_dim": cfg.encoder_ffn_dim, "output_dim": cfg.output_dim}}))
    x = synth(np.random.default_rng(5), 16000 * 3 + 123)

    seen = {}

    def fake_decoder(audio_embeddings, n_audio_tokens, language, max_tokens, **sampling):
        seen["n"] = n_audio_tokens
        seen["shape"] = audio_embeddings.shape
        seen["language"] = language
        return " hello "

    with Qwen3ASR.from_pretrained(tmp_path, decoder_backend=fake_decoder) as model:
        emb = np.array(model.encode(x))
        # same weights (rounded to bf16 by the checkpoint) loaded directly
        ref_enc = AudioEncoder(cfg)
        ref_enc.load_weights({k: bf16_round(v) for k, v in params.items()})
        assert np.array_equal(emb[0], np.array(ref_enc.encode_audio_batch([x])[0]))
        ref_enc.close()
        # ... and against the ORACLE run on the checkpoint's own (bf16-rounded) tensors: what the reference computes after
        # load_encoder_weights (encoder.py:330-359) on this file, fp32 activations
        oracle = encoder_torch.encoder_forward({k: bf16_round(v) for k, v in params.items()}, cfg, mel_np.log_mel_spectrogram_fast(x))
        assert rel_err(emb[0], oracle) <= EMB_TOL
        assert emb.shape == (1, enc.num_tokens(len(x) // 160), cfg.output_dim)
        r = model.transcribe(x, language="de")
        assert isinstance(r, TranscriptionResult) and r.text == "hello" and r.language == "German" and abs(r.duration - len(x) / 16000) < 1e-9
        assert seen["n"] == emb.shape[1] and seen["shape"] == (emb.shape[1], cfg.output_dim)
        assert model.transcribe(np.zeros(0, dtype=np.float32)) == TranscriptionResult("", "Unknown", 0.0)  # model.py:303-304
        with pytest.raises(ValueError):
            model.transcribe(np.zeros((2, 16000), dtype=np.float32))  # model.py:298-301
        # long audio: split at low-energy boundaries, every segment encoded in one varlen batch (model.py:382-447)
        long = np.concatenate([x, np.zeros(8000, dtype=np.float32), x])
        r = model.transcribe(long, chunk_duration=3.5)
        assert r.text == "hello hello"
        model.warm_up()
    with Qwen3ASR.from_pretrained(tmp_path) as model:
        with pytest.raises(NotImplementedError):
            model.transcribe(x)


def test_full_arch_config4_size_window_independence(full):
    """BASELINE config 4 size (one 20-minute utterance, 120 000 frames, 15 600 tokens, 150 attention windows): shape and
    finiteness, and the size-independent property of the block-diagonal attention (encoder.py:297-311): given the same
    log-mel, the tokens of the first k windows do not depend on anything after them -- bit for bit."""
    import torch

    from qwen3_asr_mlx_b200 import log_mel_spectrogram

    cfg, params, enc = full
    g = torch.Generator(device="cuda").manual_seed(4)
    x = 0.1 * torch.randn(1200 * 16000, device="cuda", generator=g)
    mel = log_mel_spectrogram(x)                       # (128, 120000) on the device, one utterance-wide max (audio.py:275)
    assert mel.shape == (128, 120000)
    whole = enc(mel).tensor[0]
    assert tuple(whole.shape) == (15600, 2048) and bool(torch.isfinite(whole).all())
    prefix = enc(mel.tensor[:, : 30 * 800].contiguous()).tensor[0]   # 30 windows of 800 frames = 3120 tokens
    assert tuple(prefix.shape) == (3120, 2048)
    assert torch.equal(prefix, whole[:3120])
    middle = enc(mel.tensor[:, 40 * 800: 45 * 800].contiguous()).tensor[0]  # windows restart with the chunk grid: any 800-frame-aligned cut
    assert torch.equal(middle, whole[40 * 104: 45 * 104])


def test_full_arch_config3_style_order_invariance(full):
    """BASELINE config 3 style (mixed 1-30 s utterances, varlen-packed): an utterance's embeddings do not depend on its
    position in the batch or on its neighbours -- bit for bit -- and the token counts follow the reference's rule."""
    cfg, params, enc = full
    rng = np.random.default_rng(20261018)
    lengths = [int(n) for n in rng.integers(16000, 480001, size=48)]
    xs = [synth(np.random.default_rng(100 + i), n) for i, n in enumerate(lengths)]
    emb, toffs = enc.encode_audio_batch(xs)
    e = np.array(emb)
    assert list(np.diff(toffs)) == [enc.num_tokens(n // 160) for n in lengths]
    perm = rng.permutation(len(xs))
    emb2, toffs2 = enc.encode_audio_batch([xs[i] for i in perm])
    e2 = np.array(emb2)
    for pos, i in enumerate(perm):
        assert np.array_equal(e2[int(toffs2[pos]): int(toffs2[pos + 1])], e[int(toffs[i]): int(toffs[i + 1])]), i
    assert np.isfinite(e).all()


def test_poisoned_utterance_stays_isolated(small):
    """A NaN utterance poisons ITSELF only (the reference encodes utterances one at a time, model.py:239-250).  Attention tiles
    over-read neighbouring rows of the packed batch by design; masked keys must contribute exactly 0 even when those rows are
    NaN, and NaN rows left in the workspace by a poisoned call must not leak into a later, smaller call."""
    cfg, params, enc = small
    rng = np.random.default_rng(77)
    # lengths chosen so that windows end at every residue mod 16 and the clean utterance's last window is short
    lens = [16000 * 3 + 1234, 16000 * 9 + 4321, 160 * 57, 16000 * 11, 16000 * 2 + 99]
    clean = [synth(rng, n) for n in lens]
    solo = [np.array(enc.encode_audio_batch([x])[0]) for x in clean]
    for bad_at in range(len(lens)):
        xs = [x.copy() for x in clean]
        xs[bad_at][len(xs[bad_at]) // 2] = np.nan
        emb, toffs = enc.encode_audio_batch(xs)
        e = np.array(emb)
        for u in range(len(lens)):
            got = e[int(toffs[u]): int(toffs[u + 1])]
            if u == bad_at:
                assert np.isnan(got).all()
            else:
                assert np.array_equal(got, solo[u]), (bad_at, u)
        # a clean, smaller call right after the poisoned one (stale NaN rows beyond its last token)
        for u in (2, 4):
            assert np.array_equal(np.array(enc.encode_audio_batch([clean[u]])[0]), solo[u])
    # +-inf samples behave the same way
    xs = [x.copy() for x in clean]
    xs[1][100] = np.inf
    e = np.array(enc.encode_audio_batch(xs)[0])
    toffs = enc.encode_audio_batch(xs)[1]
    assert np.array_equal(e[: int(toffs[1])], solo[0]) and np.array_equal(e[int(toffs[2]): int(toffs[3])], solo[2])


def test_weight_shapes_are_validated(small):
    """qasr_set_weight is strict like the reference's model.load_weights (encoder.py:358): a PyTorch-layout conv weight
    (O, I, kH, kW), a transposed Linear, a wrong rank or an unknown name is a ValueError, not silently wrong embeddings."""
    from qwen3_asr_mlx_b200 import AudioEncoder

    cfg, params, _ = small
    for name, bad in [
        ("conv2d2.weight", np.ascontiguousarray(params["conv2d2.weight"].transpose(0, 3, 1, 2))),  # (480,480,3,3)
        ("conv2d1.weight", params["conv2d1.weight"].reshape(480, 1, 3, 3)),
        ("layers.0.fc1.weight", np.ascontiguousarray(params["layers.0.fc1.weight"].T)),
        ("proj2.bias", params["proj2.bias"][None]),
        ("layers.2.fc1.bias", params["layers.0.fc1.bias"]),       # layer index beyond encoder_layers
        ("conv_out.bias", np.zeros(cfg.d_model, np.float32)),      # conv_out has no bias (encoder.py:174-178)
    ]:
        enc = AudioEncoder(cfg)
        p = dict(params)
        p[name] = bad
        with pytest.raises(ValueError):
            enc.load_weights(p)
        enc.close()


def test_one_pass_mel_on_the_fused_path_is_bit_identical(small, monkeypatch):
    """Waveform -> embeddings: the mel scratch keeps the raw log10 mel and conv1 applies max(x, utt_max - 8), (x + 4) / 4
    while staging its input (no normalise pass).  Bit-identical to (a) the two-pass variant and (b) mel followed by encoder
    through the separate entry points, for ragged batches including a NaN utterance and a two-lane split."""
    from qwen3_asr_mlx_b200 import AudioEncoder, log_mel_spectrogram_batch

    cfg, params, enc = small
    rng = np.random.default_rng(5)
    xs = [synth(rng, int(n)) for n in (160, 16000 * 3 + 77, 16000 * 12, 4000, 16000 * 9 + 4321)]
    xs[3] = (xs[3] * 1e-4).astype(np.float32)                     # a quiet utterance: clamp floor well above most bins
    fused = np.array(enc.encode_audio_batch(xs)[0])
    mel, foffs = log_mel_spectrogram_batch(xs)
    mels = [mel.tensor[128 * int(a): 128 * int(b)].view(128, -1) for a, b in zip(foffs[:-1], foffs[1:])]
    separate = np.array(enc.encode_batch(mels)[0])
    assert np.array_equal(fused, separate)
    monkeypatch.setenv("QASR_MEL_ONE_PASS", "0")
    two_pass = AudioEncoder(cfg)
    two_pass.load_weights(params)
    assert np.array_equal(np.array(two_pass.encode_audio_batch(xs)[0]), fused)
    l_two = two_pass.stats()["kernel_launches"]
    two_pass.encode_audio_batch(xs)
    l_two = two_pass.stats()["kernel_launches"] - l_two
    two_pass.close()
    l_one = enc.stats()["kernel_launches"]
    enc.encode_audio_batch(xs)
    assert enc.stats()["kernel_launches"] - l_one == l_two - 1   # the normalise launch is gone
    monkeypatch.delenv("QASR_MEL_ONE_PASS")
    monkeypatch.setenv("QASR_LANES", "2")
    monkeypatch.setenv("QASR_LANE_MIN_CHUNKS", "1")
    lanes = AudioEncoder(cfg)
    lanes.load_weights(params)
    assert np.array_equal(np.array(lanes.encode_audio_batch(xs)[0]), fused)   # lane 1 indexes the maxima with its utterance base
    lanes.close()


def test_gelu_approximation_error_over_every_bf16_input():
    """The kernels evaluate GELU as h (1 + tanh(h (2a + 8b h^2 + 32c h^4))), h = x / 2, with tanh.approx -- a deliberate
    deviation from the reference's exact-erf nn.gelu (encoder.py:118,273-275,320).  Bound it over EVERY finite bf16 input in
    [-40, 40] (the operands of those GEMMs' epilogues are fp32 sums, but their outputs are rounded to bf16, so what matters
    is the error relative to one bf16 ulp of the result) and over a dense fp32 grid."""
    import math

    from qwen3_asr_mlx_b200 import _lib

    bits = np.arange(0, 1 << 16, dtype=np.uint32)
    x16 = (bits << 16).view(np.float32)
    x16 = x16[np.isfinite(x16) & (np.abs(x16) <= 40.0)]
    grid = np.linspace(-8.0, 8.0, 400001, dtype=np.float32)
    x = np.ascontiguousarray(np.concatenate([x16, grid]))
    out = np.empty_like(x)
    lib = _lib.load()
    f32 = ctypes.POINTER(ctypes.c_float)
    _lib.check(lib.qasr_test_gelu(0, x.ctypes.data_as(f32), len(x), out.ctypes.data_as(f32)))
    x64 = x.astype(np.float64)
    exact = 0.5 * x64 * (1.0 + np.vectorize(math.erf)(x64 / math.sqrt(2.0)))
    err = np.abs(out.astype(np.float64) - exact)
    ulp = np.maximum(np.abs(exact), 2.0 ** -126) * 2.0 ** -8   # one bf16 ulp of the result (every call site rounds to bf16)
    table = {t: float((err[np.abs(exact) >= t] / ulp[np.abs(exact) >= t]).max()) for t in (1e-3, 1e-2, 0.05, 0.1, 0.5, 1.0)}
    print(f"gelu_from_half vs exact erf-GELU: max abs err {err.max():.3e} at x = {x[err.argmax()]:.4f}; max err in bf16 ulps of the result, by |gelu| >= t: {table}")
    # measured on B200: 3.0e-5 absolute (at x = 1.28); <= 0.28 bf16 ulp of the result wherever |gelu| >= 0.01; in the negative
    # tail (|gelu| < 0.01) the absolute error stays <= 3e-5 while the result itself shrinks, up to 5 ulps of a 1e-3 result
    assert err.max() <= 4e-5
    assert table[1e-2] <= 0.35 and table[0.1] <= 0.08 and table[1e-3] <= 6.0
    assert out[x == 0.0].tolist() == [0.0] * int((x == 0.0).sum()) and np.all(out[x >= 6.0] == x[x >= 6.0]) and np.all(np.abs(out[x <= -6.0]) <= 1e-6)


def test_outputs_stay_inside_their_buffers(small):
    """compute-sanitizer is closed on this pool, so every caller-visible output is checked with canary borders instead: the
    kernels over-READ neighbouring rows by design (128-row TMA tiles, see attention_sm100.cuh) but must never WRITE outside
    [out, out + rows * dim), whatever the (ragged, non-tile-multiple) shape.  Embeddings fp32 / bf16 via qasr_encode_audio and
    qasr_encode, the log-mel via qasr_mel, and the handle's workspace is exercised at growing and shrinking sizes."""
    import torch

    from qwen3_asr_mlx_b200 import log_mel_spectrogram_packed

    cfg, params, enc = small
    rng = np.random.default_rng(123)
    pad = 4104  # elements of canary on each side: not a multiple of any tile, but keeps the 16-byte alignment TMA stores need
    for lens in ([16000 * 7 + 13], [160, 16000 * 2 + 1, 16000 * 9 + 4321], [16000 * 30, 4000, 16000 * 12 + 7, 161]):
        xs = [synth(rng, n) for n in lens]
        soffs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        audio = torch.from_numpy(np.concatenate(xs)).cuda()
        n_tok = sum(enc.num_tokens(n // 160) for n in lens)
        for dt, name, canary in ((torch.float32, "float32", 12345.0), (torch.bfloat16, "bfloat16", 12352.0)):
            big = torch.full((2 * pad + n_tok * cfg.output_dim,), canary, dtype=dt, device="cuda")
            out = big[pad: pad + n_tok * cfg.output_dim].view(n_tok, cfg.output_dim)
            for _ in range(3):  # eager, captured, replayed
                enc.encode_packed_audio(audio, soffs, out_dtype=name, out=out)
            torch.cuda.synchronize()
            assert bool((big[:pad] == canary).all()) and bool((big[pad + n_tok * cfg.output_dim:] == canary).all()), (lens, name)
            assert bool(torch.isfinite(out.float()).all()) and not bool((out == canary).all(dim=1).any())
        with pytest.raises(ValueError):  # outputs must be 16-byte aligned (TMA tile stores); misalignment is an argument error
            off = torch.empty((n_tok * cfg.output_dim + 1,), dtype=torch.float32, device="cuda")
            enc.encode_packed_audio(audio, soffs, out=off[1:].view(n_tok, cfg.output_dim))
        # mel output, packed (128, T_u) blocks
        frames = sum(n // 160 for n in lens)
        ref_mel, _ = log_mel_spectrogram_packed(audio, soffs)
        h = enc._handle
        bigm = torch.full((2 * pad + 128 * frames,), 777.0, device="cuda")
        melv = bigm[pad: pad + 128 * frames]
        h.check(h.lib.qasr_mel(h.ptr, ctypes.c_void_p(audio.data_ptr()), soffs.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), len(lens),
                               ctypes.c_void_p(melv.data_ptr()), h.stream_ptr()))
        torch.cuda.synchronize()
        assert bool((bigm[:pad] == 777.0).all()) and bool((bigm[pad + 128 * frames:] == 777.0).all())
        assert bool(torch.equal(melv, ref_mel.tensor))


@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
def test_hidden_plus_row_block_projector_equals_one_call(small, dtype):
    """qasr_encode_audio_hidden + qasr_project_rows (the launcher's tail-block path): any blocking of the projector rows,
    in any order, gives the bits of one qasr_encode_audio call -- eager, captured and replayed."""
    import torch

    from qwen3_asr_mlx_b200 import _lib

    cfg, params, enc = small
    rng = np.random.default_rng(91)
    lens = [16000 * 11 + 123, 16000 * 3, 4000, 16000 * 29 + 1]
    x = np.concatenate([synth(rng, n) for n in lens])
    soffs = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=soffs[1:])
    audio = torch.from_numpy(x).cuda()
    tdt = torch.bfloat16 if dtype == "bfloat16" else torch.float32
    ref, toffs = enc.encode_packed_audio(audio, soffs, out_dtype=dtype)
    ref = ref.tensor.clone()
    n = int(toffs[-1])
    for blocks in ([n], [256, n - 256], [100, 7, n - 107], [n - 1, 1]):
        for _ in range(3):  # eager, capture, replay of the hidden-state call
            n_hidden, toffs2 = enc.encode_packed_audio_hidden(audio, soffs)
            assert n_hidden == n and np.array_equal(toffs2, toffs)
            out = torch.full((n, cfg.output_dim), float("nan"), dtype=tdt, device="cuda")
            starts = np.concatenate([[0], np.cumsum(blocks)[:-1]])
            for r0, nr in reversed(list(zip(starts, blocks))):  # order does not matter
                enc.project_rows(int(r0), out[int(r0): int(r0) + int(nr)])
            assert torch.equal(out.view(torch.int16 if tdt == torch.bfloat16 else torch.int32),
                               ref.view(torch.int16 if tdt == torch.bfloat16 else torch.int32))
    with pytest.raises(ValueError):
        enc.project_rows(n - 3, torch.empty((8, cfg.output_dim), dtype=tdt, device="cuda"))  # rows past the call
    enc.encode_packed_audio(audio, soffs, out_dtype=dtype)  # an ordinary call drops the hidden state
    with pytest.raises(_lib.QasrError):
        enc.project_rows(0, torch.empty((8, cfg.output_dim), dtype=tdt, device="cuda"))


def test_many_call_shapes_keep_their_graphs(small):
    """A ragged job cycles through one call shape per sub-batch (config 3: 26 of them): every shape keeps its captured
    graph (cache of 64) instead of being re-captured on every pass, and results stay those of the eager pass."""
    import torch

    cfg, params, enc = small
    rng = np.random.default_rng(92)
    shapes = [16000 + 160 * k for k in range(20)]
    audio = [torch.from_numpy(synth(rng, n)).cuda() for n in shapes]
    outs = [torch.empty((enc.num_tokens(n // 160), cfg.output_dim), dtype=torch.float32, device="cuda") for n in shapes]
    first = None
    for p in range(4):
        for a, o, n in zip(audio, outs, shapes):
            enc.encode_packed_audio(a, np.array([0, n], dtype=np.int64), out=o)
        got = [o.cpu().numpy().copy() for o in outs]
        if first is None:
            first = got
        assert all(np.array_equal(g, f) for g, f in zip(got, first))
    torch.cuda.synchronize()
