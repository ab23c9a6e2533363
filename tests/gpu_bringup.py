"""GPU bring-up diagnostics (not a pytest file): each stage runs in its own subprocess with a
timeout so that a CUDA fault in one stage cannot poison the others.

    python tests/gpu_bringup.py            # all stages
    python tests/gpu_bringup.py gemm mel   # selected stages
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def bf16_round(a: np.ndarray) -> np.ndarray:
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) >> 16
    return (u.astype(np.uint32) << 16).view(np.float32)


def to_bf16_bits(a: np.ndarray) -> np.ndarray:
    return (bf16_round(a).view(np.uint32) >> 16).astype(np.uint16)


def stage_gemm_pair():
    """cta_group::2 kernel (mode + 16): store / gelu / residual epilogues."""
    from qwen3_asr_mlx_b200 import _lib

    lib = _lib.load()
    rng = np.random.default_rng(0)
    for (M, N, K) in [(256, 256, 64), (128, 256, 128), (300, 1024, 1024), (1000, 3072, 1024), (2000, 1024, 4096), (13, 256, 7680)]:
        a = bf16_round(rng.standard_normal((M, K)).astype(np.float32))
        w = bf16_round((rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32))
        bias = rng.standard_normal(N).astype(np.float32)
        ab, wb = to_bf16_bits(a), to_bf16_bits(w)
        ref = a.astype(np.float64) @ w.astype(np.float64).T + bias
        for mode in (16, 18):
            out = rng.standard_normal((M, N)).astype(np.float32) if mode == 18 else np.zeros((M, N), dtype=np.float32)
            base = out.copy()
            rc = lib.qasr_test_gemm(0, ab.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)), wb.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)),
                                    bias.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), M, N, K, mode,
                                    out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
            if rc != 0:
                print(f"gemm_pair {M}x{N}x{K} mode {mode}: rc={rc} {_lib.last_error()}")
                continue
            want = ref + base if mode == 18 else ref
            err = np.abs(out - want).max()
            print(f"gemm_pair {M}x{N}x{K} mode {mode}: max_abs_err={err:.3e} {'OK' if err < 2e-3 else 'FAIL'}")
            if err >= 2e-3:
                bad = np.argwhere(np.abs(out - want) > 2e-3)
                print("   rows bad:", np.unique(bad[:, 0])[:12].tolist(), "cols bad:", np.unique(bad[:, 1])[:12].tolist(), "n_bad", len(bad))


def stage_gemm():
    from qwen3_asr_mlx_b200 import _lib

    lib = _lib.load()
    rng = np.random.default_rng(0)
    shapes = [(128, 256, 64), (128, 256, 128), (128, 256, 512), (256, 512, 256), (300, 1024, 1024), (13, 256, 7680),
              (1000, 3072, 1024), (2000, 1024, 4096), (130, 2048, 1024)]
    for (M, N, K) in shapes:
        a = bf16_round(rng.standard_normal((M, K)).astype(np.float32))
        w = bf16_round((rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32))
        bias = rng.standard_normal(N).astype(np.float32)
        out = np.zeros((M, N), dtype=np.float32)
        ab, wb = to_bf16_bits(a), to_bf16_bits(w)
        for mode in (0, 1):
            rc = lib.qasr_test_gemm(0, ab.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)), wb.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)),
                                    bias.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), M, N, K, mode,
                                    out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
            if rc != 0:
                print(f"gemm {M}x{N}x{K} mode {mode}: rc={rc} {_lib.last_error()}")
                continue
            ref = a.astype(np.float64) @ w.astype(np.float64).T + bias
            if mode == 1:
                from scipy.special import erf
                ref = 0.5 * ref * (1 + erf(ref / np.sqrt(2)))
            err = np.abs(out - ref).max()
            print(f"gemm {M}x{N}x{K} mode {mode}: max_abs_err={err:.3e} ref_max={np.abs(ref).max():.3f} {'OK' if err < 2e-3 else 'FAIL'}")
            if err >= 2e-3:
                bad = np.argwhere(np.abs(out - ref) > 2e-3)
                print("   first bad:", bad[:5].tolist(), "rows bad:", np.unique(bad[:, 0])[:10].tolist(), "cols bad:", np.unique(bad[:, 1])[:10].tolist(),
                      "n_bad", len(bad))


def synth(rng, n):
    t = np.arange(n) / 16000.0
    x = 0.1 * rng.standard_normal(n)
    for _ in range(3):
        x += 0.3 * np.sin(2 * np.pi * rng.uniform(100, 4000) * t + rng.uniform(0, 6.28)) * (0.5 + 0.5 * np.sin(2 * np.pi * rng.uniform(0.1, 1.0) * t))
    return np.clip(x, -1, 1).astype(np.float32)


def stage_mel():
    from oracle import mel_np
    from qwen3_asr_mlx_b200 import audio

    rng = np.random.default_rng(1)
    fb = np.empty((128, 201), dtype=np.float32)
    from qwen3_asr_mlx_b200 import _lib
    lib = _lib.load()
    lib.qasr_mel_filterbank(fb.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    ref_fb = mel_np.mel_filterbank()
    print("filterbank max abs diff vs oracle:", np.abs(fb - ref_fb).max(), "max rel:", (np.abs(fb - ref_fb) / np.maximum(ref_fb, 1e-30)).max())
    win = np.empty(400, dtype=np.float32)
    lib.qasr_hann_window(win.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    print("hann max abs diff:", np.abs(win - np.hanning(400).astype(np.float32)).max())
    for n in (160, 199, 201, 400, 5000, 16000, 160000, 480000, 16000 * 7 + 123):
        x = synth(rng, n)
        got = np.array(audio.log_mel_spectrogram(x))
        ref = mel_np.log_mel_spectrogram_fast(x)
        err = np.abs(got - ref).max()
        print(f"mel N={n}: shape {got.shape} max_abs_err={err:.3e} {'OK' if err <= 1e-4 else 'FAIL'}")
    for name, x in (("silence", np.zeros(16000, np.float32)), ("tone", np.sin(2 * np.pi * 440 * np.arange(16000) / 16000).astype(np.float32))):
        got = np.array(audio.log_mel_spectrogram(x))
        ref = mel_np.log_mel_spectrogram_fast(x)
        print(f"mel {name}: max_abs_err={np.abs(got - ref).max():.3e} min {got.min():.4f} max {got.max():.4f}")
    xs = [synth(rng, int(n)) for n in rng.integers(16000, 200000, size=9)]
    mel, foffs = audio.log_mel_spectrogram_batch(xs)
    m = np.array(mel)
    worst = 0.0
    for u, x in enumerate(xs):
        T = int(foffs[u + 1] - foffs[u])
        blk = m[128 * int(foffs[u]): 128 * int(foffs[u + 1])].reshape(128, T)
        worst = max(worst, np.abs(blk - mel_np.log_mel_spectrogram_fast(x)).max())
    print(f"mel batch of 9 ragged: max_abs_err={worst:.3e} {'OK' if worst <= 1e-4 else 'FAIL'}")


def rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b), 1e-30))


def stage_enc_small():
    from oracle import encoder_torch, mel_np
    from qwen3_asr_mlx_b200 import weights
    from qwen3_asr_mlx_b200.config import AudioEncoderConfig
    from qwen3_asr_mlx_b200.encoder import AudioEncoder

    cfg = AudioEncoderConfig(d_model=256, encoder_layers=2, encoder_attention_heads=4, encoder_ffn_dim=512, output_dim=256)
    P = weights.random_init(cfg, seed=1, exercise_all=True)
    enc = AudioEncoder(cfg)
    enc.load_weights(P)
    enc.set_debug(True)
    rng = np.random.default_rng(0)
    for n in (16000, 16000 * 4 + 7000, 16000 * 20 + 333):
        x = synth(rng, n)
        mel = mel_np.log_mel_spectrogram_fast(x)
        ref, inter = encoder_torch.encoder_forward(P, cfg, mel, return_intermediates=True)
        out = np.array(enc(mel))[0]
        nt = ref.shape[0]
        print(f"enc_small N={n}: tokens {out.shape} vs {ref.shape}")
        for k in ("stem", "layer0", "hidden"):
            got = enc.debug_read(k, nt)
            print(f"   {k:7s} rel_err={rel(got, inter[k]):.3e} max_abs={np.abs(got - inter[k]).max():.3e} ref_max={np.abs(inter[k]).max():.3f}")
        r = rel(out, ref)
        print(f"   output  rel_err={r:.3e} {'OK' if r <= 2e-2 else 'FAIL'}")


def stage_enc_full():
    import torch
    from oracle import encoder_torch, mel_np
    from qwen3_asr_mlx_b200 import weights
    from qwen3_asr_mlx_b200.config import AudioEncoderConfig
    from qwen3_asr_mlx_b200.encoder import AudioEncoder

    cfg = AudioEncoderConfig()
    t0 = time.time()
    P = weights.random_init(cfg, seed=1234)
    print(f"random_init {time.time() - t0:.1f}s")
    enc = AudioEncoder(cfg)
    t0 = time.time()
    enc.load_weights(P)
    print(f"load_weights {time.time() - t0:.1f}s")
    enc.set_debug(True)
    rng = np.random.default_rng(0)
    x = synth(rng, 160000)
    mel = mel_np.log_mel_spectrogram_fast(x)
    t0 = time.time()
    ref, inter = encoder_torch.encoder_forward(P, cfg, mel, return_intermediates=True)
    print(f"oracle 10 s utterance: {time.time() - t0:.1f}s on {torch.get_num_threads()} threads")
    out = np.array(enc(mel))[0]
    for k in ("stem", "layer0", "hidden"):
        got = enc.debug_read(k, ref.shape[0])
        print(f"   {k:7s} rel_err={rel(got, inter[k]):.3e}")
    r = rel(out, ref)
    print(f"enc_full cfg1 (10 s): out {out.shape} rel_err={r:.3e} {'OK' if r <= 2e-2 else 'FAIL'}")
    # end-to-end from audio
    emb, toffs = enc.encode_audio_batch([x])
    print(f"   from audio: rel_err={rel(np.array(emb), ref):.3e}")
    enc.set_debug(False)
    # config 2 timing: 64 x 30 s
    xs = np.concatenate([synth(rng, 480000) for _ in range(4)] * 16)
    soffs = np.arange(65, dtype=np.int64) * 480000
    audio_dev = torch.from_numpy(xs).cuda()
    enc.reserve(64 * 3000, 64)
    for it in range(3):
        emb, toffs = enc.encode_packed_audio(audio_dev, soffs)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    iters = 5
    for it in range(iters):
        emb, toffs = enc.encode_packed_audio(audio_dev, soffs)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / iters
    print(f"cfg2 64x30s: {ms:.2f} ms/step -> {1920.0 / (ms / 1e3):.0f} audio-s/s, {23.947e12 / (ms / 1e3) / 1e12:.1f} TFLOP/s algorithmic; tokens {emb.shape}")
    print("stats", enc.stats())


STAGES = {"gemm": stage_gemm, "gemm_pair": stage_gemm_pair, "mel": stage_mel, "enc_small": stage_enc_small, "enc_full": stage_enc_full}

if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--stage":
        try:
            STAGES[sys.argv[2]]()
        except Exception:
            traceback.print_exc()
            sys.exit(1)
        sys.exit(0)
    names = sys.argv[1:] or list(STAGES)
    for name in names:
        print(f"===== stage {name}", flush=True)
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--stage", name], timeout=420, cwd=ROOT)
            print(f"===== stage {name} exit {r.returncode} in {time.time() - t0:.1f}s", flush=True)
        except subprocess.TimeoutExpired:
            print(f"===== stage {name} TIMEOUT", flush=True)
