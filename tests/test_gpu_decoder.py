"""GPU parity of the decoder prefill (through the C ABI, include/qasr_decoder.h) against the oracle.
Tolerances (bf16 tensor-core kernels with fp32 accumulation vs the fp32 oracle; the north_star gives 2e-2 for the encoder
embeddings, the same bar is used here): relative Frobenius error <= 2e-2 for the residual stream, the cached keys / values
and the logits."""
import numpy as np
import pytest

from helpers import EMB_TOL, rel_err
from oracle import decoder_torch

pytestmark = pytest.mark.gpu


def _small_cfg():
    from qwen3_asr_mlx_b200.config import TextDecoderConfig

    return TextDecoderConfig(hidden_size=256, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, intermediate_size=512, vocab_size=1024)


@pytest.fixture(scope="module")
def small():
    from qwen3_asr_mlx_b200 import decoder as dec

    cfg = _small_cfg()
    params = dec.random_init(cfg, seed=5, exercise_all=True)
    d = dec.TextDecoder(cfg)
    d.load_weights(params)
    yield cfg, params, d
    d.close()


def _emb(seed, n, hidden):
    import torch

    return torch.randn(n, hidden, generator=torch.Generator().manual_seed(seed))


def test_ragged_batch_parity_and_cache_layout(small):
    cfg, params, d = small
    lens = [1, 63, 64, 65, 130, 407, 7]
    offs = np.concatenate([[0], np.cumsum(lens)])
    emb = _emb(1, int(offs[-1]), cfg.hidden_size)
    last, cache, full, hid = d.prefill(emb.cuda(), offs, all_logits=True, return_hidden=True)
    last, full, hid = np.array(last), np.array(full), np.array(hid)
    assert last.shape == (len(lens), cfg.vocab_size) and full.shape == (int(offs[-1]), cfg.vocab_size)
    assert cache.offset == lens
    refs = decoder_torch.decoder_prefill_batch(params, cfg, emb, offs)
    for u, r in enumerate(refs):
        a, b = int(offs[u]), int(offs[u + 1])
        assert rel_err(hid[a:b], r["hidden"]) <= EMB_TOL, u
        assert rel_err(full[a:b], r["logits"]) <= EMB_TOL, u
        assert np.array_equal(last[u], full[b - 1])  # the last-position logits are the full logits' last row
        for layer in range(cfg.num_hidden_layers):
            k, v = cache.layer(layer, u)  # (1, n_kv, T, head_dim), the reference's shapes (decoder.py:43-46)
            assert tuple(k.shape) == (1, cfg.num_key_value_heads, b - a, cfg.head_dim)
            assert rel_err(k[0].float().cpu().numpy(), r["keys"][layer]) <= EMB_TOL, (u, layer)
            assert rel_err(v[0].float().cpu().numpy(), r["values"][layer]) <= EMB_TOL, (u, layer)


def test_batch_equals_loop_of_singles(small):
    cfg, params, d = small
    lens = [50, 129, 3]
    offs = np.concatenate([[0], np.cumsum(lens)])
    emb = _emb(2, int(offs[-1]), cfg.hidden_size).cuda()
    last, cache, full = d.prefill(emb, offs, all_logits=True)
    for u in range(len(lens)):
        a, b = int(offs[u]), int(offs[u + 1])
        l1, c1, f1 = d.prefill(emb[a:b], all_logits=True)
        assert np.array_equal(np.array(f1), np.array(full)[a:b])  # batching must not change results
        assert np.array_equal(np.array(l1)[0], np.array(last)[u])
        assert bool((c1.keys == cache.keys[:, a:b]).all()) and bool((c1.values == cache.values[:, a:b]).all())


def test_reference_shaped_call_and_inputs(small):
    import torch

    cfg, params, d = small
    emb = _emb(3, 40, cfg.hidden_size).cuda()
    logits = d(emb[None], cache=None, is_embeds=True)  # decoder.py:223-253: (1, T, hidden) -> (1, T, vocab)
    assert logits.shape == (1, 40, cfg.vocab_size)
    bf = d.prefill(emb.bfloat16(), return_cache=False)[0]  # bf16 embeddings (the dtype prepare_inputs produces)
    ref = decoder_torch.decoder_prefill(params, cfg, emb.bfloat16().float().cpu())["logits"][-1]
    assert rel_err(np.array(bf)[0], ref) <= EMB_TOL
    table = d.embed_tokens  # the library's bf16 copy of embed_tokens.weight
    assert tuple(table.shape) == (cfg.vocab_size, cfg.hidden_size) and table.dtype == torch.bfloat16
    assert torch.equal(table.float().cpu(), params["embed_tokens.weight"].bfloat16().float())
    ids = torch.tensor([[5, 17, 1000, 3]])
    by_ids = np.array(d(ids))  # token ids in, like the reference's default is_embeds=False
    by_emb = np.array(d(table[ids[0]][None], is_embeds=True))
    assert np.array_equal(by_ids, by_emb)
    with pytest.raises(ValueError):
        d.prefill(emb, [0, 10])  # offsets must end at n
    with pytest.raises(ValueError):
        d.prefill(emb[:, :100])
    with pytest.raises(ValueError):
        d.prefill(emb.cpu())


def test_weight_errors():
    from qwen3_asr_mlx_b200 import _lib
    from qwen3_asr_mlx_b200 import decoder as dec

    cfg = _small_cfg()
    params = dec.random_init(cfg, seed=5)
    d = dec.TextDecoder(cfg)
    with pytest.raises(ValueError):
        d.load_weights({"layers.0.mlp.bogus.weight": np.zeros((2, 2), np.float32)})
    with pytest.raises(ValueError):
        d.load_weights({"norm.weight": np.zeros(17, np.float32)})
    part = dict(params)
    del part["layers.1.mlp.up_proj.weight"]
    with pytest.raises(_lib.QasrError):
        d.load_weights(part)  # finalize reports the missing parameter
    d.close()


def test_weight_sources_agree(small, tmp_path):
    """numpy fp32, CPU bf16, CUDA fp32 and a safetensors checkpoint with the reference's `model.` prefix give the same decoder."""
    import torch
    from safetensors.torch import save_file

    from qwen3_asr_mlx_b200 import decoder as dec

    cfg, params, d = small
    emb = _emb(4, 33, cfg.hidden_size).cuda()
    want = np.array(d.prefill(emb, return_cache=False)[0])
    variants = {
        "numpy": {k: v.numpy() for k, v in params.items()},
        "cuda_fp32": {k: v.cuda() for k, v in params.items()},
    }
    for name, p in variants.items():
        alt = dec.TextDecoder(cfg)
        alt.load_weights(p)
        assert np.array_equal(np.array(alt.prefill(emb, return_cache=False)[0]), want), name
        alt.close()
    tensors = {"model." + k: v.to(torch.bfloat16) for k, v in params.items()}
    tensors["audio_tower.ln_post.weight"] = torch.zeros(4, dtype=torch.bfloat16)  # encoder tensors are ignored
    save_file(tensors, str(tmp_path / "model.safetensors"))
    alt = dec.TextDecoder(cfg)
    dec.load_decoder_weights(alt, tmp_path)
    # matrices are stored as bf16 either way; the checkpoint also rounds the (randomised) norm weights to bf16
    assert rel_err(np.array(alt.prefill(emb, return_cache=False)[0]), want) <= 5e-3
    alt.close()


def test_full_width_two_layers_parity():
    """1.7B widths (hidden 2048, 16 q / 8 kv heads x 128, intermediate 6144), 2 layers, reduced vocabulary; 30 s-sized prompt."""
    from qwen3_asr_mlx_b200 import decoder as dec
    from qwen3_asr_mlx_b200.config import TextDecoderConfig

    cfg = TextDecoderConfig(num_hidden_layers=2, vocab_size=4096)
    params = dec.random_init(cfg, seed=11, exercise_all=True)
    d = dec.TextDecoder(cfg)
    d.load_weights(params)
    lens = [407, 150]
    offs = np.concatenate([[0], np.cumsum(lens)])
    emb = _emb(7, int(offs[-1]), cfg.hidden_size)
    last, cache, hid = d.prefill(emb.cuda(), offs, return_hidden=True)
    refs = decoder_torch.decoder_prefill_batch(params, cfg, emb, offs)
    for u, r in enumerate(refs):
        a, b = int(offs[u]), int(offs[u + 1])
        assert rel_err(np.array(hid)[a:b], r["hidden"]) <= EMB_TOL
        assert rel_err(np.array(last)[u], r["logits"][-1]) <= EMB_TOL
        assert rel_err(cache.layer(1, u)[0][0].float().cpu().numpy(), r["keys"][1]) <= EMB_TOL
    d.close()


def test_encoder_to_prefill_pipeline():
    """The whole batched path: waveform -> mel -> encoder -> build_prompt / prepare_inputs -> decoder prefill, two utterances,
    against the oracles chained the same way (reference model.py:331-342, generate.py:266-278)."""
    import torch

    from oracle import encoder_torch, mel_np
    from qwen3_asr_mlx_b200 import AudioEncoder, AudioEncoderConfig, build_prompt, prepare_inputs, weights
    from qwen3_asr_mlx_b200 import decoder as dec
    from qwen3_asr_mlx_b200.config import TextDecoderConfig
    from qwen3_asr_mlx_b200.tokenizer import AUDIO_PAD_TOKEN_ID
    from helpers import synth

    ecfg = AudioEncoderConfig(d_model=256, encoder_layers=2, encoder_attention_heads=4, encoder_ffn_dim=512, output_dim=256)
    eparams = weights.random_init(ecfg, seed=7, exercise_all=True)
    enc = AudioEncoder(ecfg)
    enc.load_weights(eparams)
    dcfg = TextDecoderConfig(hidden_size=256, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, intermediate_size=512,
                             vocab_size=151936 // 8 * 8)  # the prompt uses the real special-token ids
    dparams = dec.random_init(dcfg, seed=5, exercise_all=True)
    d = dec.TextDecoder(dcfg)
    d.load_weights(dparams)
    rng = np.random.default_rng(12)
    xs = [synth(rng, 16000 * 3 + 500), synth(rng, 16000 * 11)]
    emb, toffs = enc.encode_audio_batch(xs)
    prompts, rows = [], []
    for u in range(2):
        n_audio = int(toffs[u + 1] - toffs[u])
        ids = build_prompt(n_audio, [22574])  # one language-name token
        assert ids.count(AUDIO_PAD_TOKEN_ID) == n_audio
        prompts.append(ids)
        rows.append(prepare_inputs(emb[int(toffs[u]): int(toffs[u + 1])], ids, d.embed_tokens).tensor[0])
    offs = np.concatenate([[0], np.cumsum([len(p) for p in prompts])])
    last, cache = d.prefill(torch.cat(rows), offs)
    table = dparams["embed_tokens.weight"].bfloat16().float()
    for u in range(2):
        ref_emb = encoder_torch.encoder_forward(eparams, ecfg, mel_np.log_mel_spectrogram_fast(xs[u]))
        x = table[torch.tensor(prompts[u])].clone()
        x[torch.tensor(prompts[u]) == AUDIO_PAD_TOKEN_ID] = torch.from_numpy(ref_emb)
        ref = decoder_torch.decoder_prefill(dparams, dcfg, x)["logits"][-1]
        assert rel_err(np.array(last)[u], ref) <= 3e-2  # two bf16 stages chained (encoder 2e-2 budget + decoder)
    enc.close()
    d.close()


def test_attention_kernels_agree_and_long_prompts(small, monkeypatch):
    """The tcgen05 causal attention (default) against the oracle on prompts that need several KV tiles with online-softmax
    rescaling (1000 rows = 8 tiles), tile-boundary lengths, and against the mma.sync kernel (QASR_DEC_ATTN_TC=0)."""
    from qwen3_asr_mlx_b200 import decoder as dec

    cfg, params, d = small
    lens = [1000, 129, 128, 127, 300, 1, 257]
    offs = np.concatenate([[0], np.cumsum(lens)])
    emb = _emb(21, int(offs[-1]), cfg.hidden_size)
    last, cache, hid = d.prefill(emb.cuda(), offs, return_hidden=True)
    hid = np.array(hid)
    refs = decoder_torch.decoder_prefill_batch(params, cfg, emb, offs)
    for u, r in enumerate(refs):
        a, b = int(offs[u]), int(offs[u + 1])
        assert rel_err(hid[a:b], r["hidden"]) <= EMB_TOL, (u, lens[u])
        assert rel_err(np.array(last)[u], r["logits"][-1]) <= EMB_TOL, (u, lens[u])
    monkeypatch.setenv("QASR_DEC_ATTN_TC", "0")
    alt = dec.TextDecoder(cfg)
    alt.load_weights(params)
    hid2 = np.array(alt.prefill(emb.cuda(), offs, return_cache=False, return_hidden=True)[2])
    assert rel_err(hid2, hid) <= 5e-3  # same math, different accumulation order / bf16 rounding points
    for u, r in enumerate(refs):
        assert rel_err(hid2[int(offs[u]): int(offs[u + 1])], r["hidden"]) <= EMB_TOL
    alt.close()


def test_prefill_batch_shell_matches_manual_pipeline():
    """Qwen3ASR.prefill_batch (one prepare_inputs gather for all prompts) == the per-utterance pipeline."""
    import torch

    from qwen3_asr_mlx_b200 import AudioEncoder, AudioEncoderConfig, Qwen3ASR, build_prompt, prepare_inputs, weights
    from qwen3_asr_mlx_b200 import decoder as dec
    from qwen3_asr_mlx_b200.config import TextDecoderConfig
    from helpers import synth

    ecfg = AudioEncoderConfig(d_model=256, encoder_layers=1, encoder_attention_heads=4, encoder_ffn_dim=512, output_dim=256)
    enc = AudioEncoder(ecfg)
    enc.load_weights(weights.random_init(ecfg, seed=7))
    dcfg = TextDecoderConfig(hidden_size=256, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, intermediate_size=512)
    d = dec.TextDecoder(dcfg)
    d.load_weights(dec.random_init(dcfg, seed=5))
    shell = Qwen3ASR(ecfg, enc, decoder=d)
    rng = np.random.default_rng(2)
    xs = [synth(rng, n) for n in (16000 * 2, 16000 * 7 + 99, 8000)]
    last, cache, poffs, toffs = shell.prefill_batch(xs, [22574])
    assert last.shape == (3, dcfg.vocab_size) and list(np.diff(poffs)) == [int(t) + 18 for t in np.diff(toffs)]
    emb, _ = enc.encode_audio_batch(xs)
    for u in range(3):
        ids = build_prompt(int(toffs[u + 1] - toffs[u]), [22574])
        x = prepare_inputs(emb[int(toffs[u]): int(toffs[u + 1])], ids, d.embed_tokens)
        single = d.prefill(x, return_cache=False)[0]
        assert np.array_equal(np.array(single)[0], np.array(last)[u])
    with pytest.raises(NotImplementedError):
        Qwen3ASR(ecfg, enc).prefill_batch(xs, [22574])
    shell.close()


def test_many_prompts_persistent_ctas_deterministic(small):
    """300 ragged prompts -> several thousand (prompt, 128-query tile, head) items, so every persistent CTA of the tcgen05
    attention walks many items and KV steps (barrier phases, TMEM reuse, the shared-memory exchange between the two
    half-row softmax groups).  Results must be identical run to run and match the oracle on sampled prompts."""
    cfg, params, d = small
    rng = np.random.default_rng(77)
    lens = [int(v) for v in rng.integers(1, 600, size=300)]
    lens[5], lens[17], lens[100] = 128, 256, 384
    offs = np.concatenate([[0], np.cumsum(lens)])
    emb = _emb(31, int(offs[-1]), cfg.hidden_size)
    emb_dev = emb.cuda()
    first = None
    for _ in range(6):
        last, cache, hid = d.prefill(emb_dev, offs, return_hidden=True)
        got = (np.array(last), np.array(hid), cache.keys.float().cpu().numpy())
        if first is None:
            first = got
        else:
            assert all(np.array_equal(a, b) for a, b in zip(first, got))
    assert np.isfinite(first[0]).all() and np.isfinite(first[1]).all()
    for u in (0, 5, 17, 100, 123, 299, int(np.argmax(lens))):
        a, b = int(offs[u]), int(offs[u + 1])
        r = decoder_torch.decoder_prefill(params, cfg, emb[a:b])
        assert rel_err(first[1][a:b], r["hidden"]) <= EMB_TOL, (u, lens[u])
        assert rel_err(first[0][u], r["logits"][-1]) <= EMB_TOL, (u, lens[u])


def test_very_long_prompt(small):
    """A 5 000-row prompt (the prompt of a ~6.4-minute single-pass utterance): 40 query tiles, up to 40 KV steps per item."""
    cfg, params, d = small
    emb = _emb(41, 5000, cfg.hidden_size)
    last, cache, hid = d.prefill(emb.cuda(), return_hidden=True)
    r = decoder_torch.decoder_prefill(params, cfg, emb)
    assert rel_err(np.array(hid), r["hidden"]) <= EMB_TOL
    assert rel_err(np.array(last)[0], r["logits"][-1]) <= EMB_TOL
    assert rel_err(cache.layer(1)[0][0].float().cpu().numpy(), r["keys"][1]) <= EMB_TOL


def test_prefix_consistency_like_cached_decode(small):
    """Reference tests/test_decoder.py::test_cached_decode_matches_full_context restated for the prefill: the logits at
    position t of a full-context pass equal the last-position logits of a prefill over the first t + 1 tokens (atol 1e-3
    there; causal masking makes them the same computation here)."""
    cfg, params, d = small
    emb = _emb(51, 300, cfg.hidden_size).cuda()
    full = np.array(d.prefill(emb, return_cache=False, all_logits=True)[2])
    for t in (0, 3, 63, 64, 127, 128, 200, 299):
        last = np.array(d.prefill(emb[: t + 1], return_cache=False)[0])[0]
        assert np.allclose(last, full[t], atol=1e-3), t


def test_poisoned_prompt_stays_isolated(small):
    """A NaN prompt between clean ones: the clean prompts' logits and KV rows are bit-identical to a clean batch (the last,
    partial KV tile of a prompt over-reads the next prompt's rows; masked keys must contribute exactly 0)."""
    cfg, params, d = small
    lens = [37, 200, 65, 129, 1]
    emb = _emb(123, sum(lens), cfg.hidden_size).cuda()
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    base, cache0 = d.prefill(emb, offs)
    base = np.array(base)
    k0, v0 = cache0.keys.float().cpu().numpy(), cache0.values.float().cpu().numpy()
    for bad in range(len(lens)):
        e2 = emb.clone()
        e2[int(offs[bad]) + lens[bad] // 2] = float("nan")
        got, cache = d.prefill(e2, offs)
        got = np.array(got)
        k1, v1 = cache.keys.float().cpu().numpy(), cache.values.float().cpu().numpy()
        for u in range(len(lens)):
            a, b = int(offs[u]), int(offs[u + 1])
            if u == bad:
                assert np.isnan(got[u]).all()
            else:
                assert np.array_equal(got[u], base[u]), (bad, u)
                assert np.array_equal(k1[:, a:b], k0[:, a:b]) and np.array_equal(v1[:, a:b], v0[:, a:b])
    # a clean call after the poisoned ones
    assert np.array_equal(np.array(d.prefill(emb, offs)[0]), base)
