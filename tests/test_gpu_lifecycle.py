"""Handle lifecycle, streams and workspace growth of libqasr (GPU): what a serving process does around the hot path.
The reference keeps one model per process and serialises transcribe() with a lock (model.py:145,239); here one handle per GPU
per worker is the unit, so handles must be independent, stream-ordered and leak-free."""
import numpy as np
import pytest

from helpers import synth

pytestmark = pytest.mark.gpu


def _cfg():
    from qwen3_asr_mlx_b200 import AudioEncoderConfig

    return AudioEncoderConfig(d_model=256, encoder_layers=2, encoder_attention_heads=4, encoder_ffn_dim=512, output_dim=256)


def test_two_handles_on_two_streams_match_serial_results():
    """Two handles fed from two non-default streams at the same time (eager, captured and replayed calls) give the bits of
    serial calls on the default stream: no hidden global state, launches are ordered on the caller's stream."""
    import torch

    from qwen3_asr_mlx_b200 import AudioEncoder, weights

    cfg = _cfg()
    params = weights.random_init(cfg, seed=21, exercise_all=True)
    encs = [AudioEncoder(cfg), AudioEncoder(cfg)]
    for e in encs:
        e.load_weights(params)
    rng = np.random.default_rng(7)
    xs = [synth(rng, 16000 * 9 + 11), synth(rng, 16000 * 14 + 500)]
    want = [np.array(encs[0].encode_audio_batch([x])[0]) for x in xs]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    dev = [torch.from_numpy(x).cuda() for x in xs]
    soffs = [np.array([0, len(x)], dtype=np.int64) for x in xs]
    outs = [torch.empty((encs[i].num_tokens(len(xs[i]) // 160), cfg.output_dim), dtype=torch.float32, device="cuda") for i in range(2)]
    torch.cuda.synchronize()
    for _ in range(4):
        for o in outs:
            o.zero_()
        torch.cuda.synchronize()
        for i in (0, 1):
            with torch.cuda.stream(streams[i]):
                encs[i].encode_packed_audio(dev[i], soffs[i], out=outs[i])
        for s in streams:
            s.synchronize()
        for i in (0, 1):
            assert np.array_equal(outs[i].cpu().numpy(), want[i])
    for e in encs:
        e.close()


def test_create_destroy_cycles_release_device_memory():
    """qasr_destroy frees everything the handle owns (weights, workspace, graphs, pinned staging): ten create / load / encode /
    close cycles leave the device's free memory where it was."""
    import torch

    from qwen3_asr_mlx_b200 import AudioEncoder, weights

    cfg = _cfg()
    params = weights.random_init(cfg, seed=22)
    x = synth(np.random.default_rng(8), 16000 * 12)

    def cycle():
        enc = AudioEncoder(cfg)
        enc.load_weights(params)
        for _ in range(3):  # eager, captured, replayed
            out = np.array(enc.encode_audio_batch([x])[0])
        assert enc.stats()["workspace_bytes"] > 0
        enc.close()
        return out

    first = cycle()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    free0, _ = torch.cuda.mem_get_info()
    for _ in range(10):
        assert np.array_equal(cycle(), first)
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    free1, _ = torch.cuda.mem_get_info()
    assert free0 - free1 < 32 * 1024 * 1024, f"device memory leaked: {free0 - free1} bytes over 10 cycles"


def test_workspace_growth_invalidates_graphs_and_keeps_results():
    """A captured call, then a much larger call (the grow-only workspace is reallocated: captured graphs hold stale pointers
    and are dropped), then the first call again: same bits as before the growth."""
    import torch

    from qwen3_asr_mlx_b200 import AudioEncoder, weights

    cfg = _cfg()
    enc = AudioEncoder(cfg)
    enc.load_weights(weights.random_init(cfg, seed=23, exercise_all=True))
    rng = np.random.default_rng(9)
    small = torch.from_numpy(synth(rng, 16000 * 5)).cuda()
    so_small = np.array([0, small.numel()], dtype=np.int64)
    out_small = torch.empty((enc.num_tokens(small.numel() // 160), cfg.output_dim), dtype=torch.float32, device="cuda")
    for _ in range(3):
        enc.encode_packed_audio(small, so_small, out=out_small)
    before = out_small.cpu().numpy().copy()
    ws0 = enc.stats()["workspace_bytes"]
    big = [synth(rng, 16000 * 30) for _ in range(12)]
    big_emb, _ = enc.encode_audio_batch(big)
    assert enc.stats()["workspace_bytes"] > ws0
    assert bool(torch.isfinite(big_emb.tensor).all().item())
    for _ in range(3):
        out_small.zero_()
        enc.encode_packed_audio(small, so_small, out=out_small)
        assert np.array_equal(out_small.cpu().numpy(), before)
    enc.close()


def test_bad_arguments_leave_the_handle_usable():
    """Every argument error comes back as the reference's exception type (ValueError) and the next valid call works."""
    import torch

    from qwen3_asr_mlx_b200 import AudioEncoder, weights

    cfg = _cfg()
    enc = AudioEncoder(cfg)
    enc.load_weights(weights.random_init(cfg, seed=24))
    x = synth(np.random.default_rng(10), 16000 * 3)
    good = np.array(enc.encode_audio_batch([x])[0])
    audio = torch.from_numpy(x).cuda()
    with pytest.raises(ValueError):
        enc.encode_packed_audio(audio, np.array([0, 100], dtype=np.int64))          # fewer than 160 samples
    with pytest.raises(ValueError):
        enc.encode_packed_audio(audio, np.array([0, 16000, 8000], dtype=np.int64))  # decreasing offsets
    with pytest.raises(ValueError):
        enc.encode_packed_audio(audio, np.array([0, len(x)], dtype=np.int64),
                                out=torch.empty((3, cfg.output_dim), dtype=torch.float32, device="cuda"))  # wrong output shape
    with pytest.raises(ValueError):
        enc.encode_audio_batch([np.zeros((2, 16000), dtype=np.float32)])            # non-1-D audio (model.py:298-301)
    assert np.array_equal(np.array(enc.encode_audio_batch([x])[0]), good)
    enc.close()
