"""Measure BASELINE.json configs 1, 3, 4, 5 on one B200 (config 2 is bench.py's line).
Writes one JSON document to gpurun_out/configs_r01.json.  Not a pytest file.

    python tests/run_configs.py
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import load_peaks, synth  # noqa: E402
from qwen3_asr_mlx_b200 import AudioEncoder, AudioEncoderConfig, launcher, weights  # noqa: E402
from qwen3_asr_mlx_b200.audio import log_mel_spectrogram_batch  # noqa: E402

SR = 16000


def timed(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    peaks = load_peaks()
    cfg = AudioEncoderConfig()
    enc = AudioEncoder(cfg, device=0)
    enc.load_weights(weights.random_init(cfg, seed=1234))
    out = {"peaks": peaks}

    # ---- config 1: one 10 s utterance (latency-bound: 130 tokens cannot fill 148 SMs)
    x = torch.from_numpy(synth(np.random.default_rng(0), 10 * SR)).cuda()
    so = np.array([0, 10 * SR], dtype=np.int64)
    emb = torch.empty((130, cfg.output_dim), dtype=torch.float32, device="cuda")
    ms = timed(lambda: enc.encode_packed_audio(x, so, out=emb), 50)
    out["config1_single_10s"] = {"ms_per_call": ms, "audio_s_per_s": 10.0 / (ms / 1e3), "tokens": 130, "note": "CUDA-graph replay, device-resident audio"}

    # ---- config 4: one 20-minute utterance, single pass (150 windows)
    x = (0.1 * torch.randn(1200 * SR, device="cuda")).contiguous()
    so = np.array([0, 1200 * SR], dtype=np.int64)
    emb = torch.empty((15600, cfg.output_dim), dtype=torch.float32, device="cuda")
    ms = timed(lambda: enc.encode_packed_audio(x, so, out=emb), 10)
    out["config4_20min_single_pass"] = {"ms_per_call": ms, "audio_s_per_s": 1200.0 / (ms / 1e3), "tokens": 15600, "tflops_algorithmic": 14974.7e9 / (ms / 1e3) / 1e12}
    del x, emb

    # ---- config 3: 4096 mixed-length utterances (1-30 s, length seed 20261018), varlen-packed, 1 GPU
    lengths = np.random.default_rng(20261018).integers(16000, 480001, size=4096)
    soffs_all = np.zeros(4097, dtype=np.int64)
    np.cumsum(lengths, out=soffs_all[1:])
    audio = (0.1 * torch.randn(int(soffs_all[-1]), device="cuda")).contiguous()  # 4 GB, device resident
    costs = [launcher.tokens_for_samples(int(n)) for n in lengths]
    order = sorted(range(4096), key=lambda i: -costs[i])  # long utterances first; each sub-batch is gathered into a packed buffer
    subs = launcher.split_by_budget(order, costs, 32768)
    total_tok = sum(costs)
    emb_all = torch.empty((total_tok, cfg.output_dim), dtype=torch.float32, device="cuda")
    tok_off = np.zeros(4097, dtype=np.int64)
    np.cumsum(costs, out=tok_off[1:])

    def run_cfg3():
        for sub in subs:
            so = np.zeros(len(sub) + 1, dtype=np.int64)
            np.cumsum([int(lengths[i]) for i in sub], out=so[1:])
            packed = torch.cat([audio[int(soffs_all[i]): int(soffs_all[i + 1])] for i in sub])
            e, t = enc.encode_packed_audio(packed, so)
            # scatter back to the original utterance order
            pos = 0
            for i in sub:
                n = costs[i]
                emb_all[int(tok_off[i]): int(tok_off[i]) + n].copy_(e.tensor[pos: pos + n])
                pos += n

    t0 = time.perf_counter()
    run_cfg3()
    torch.cuda.synchronize()
    first = time.perf_counter() - t0
    t0 = time.perf_counter()
    run_cfg3()
    torch.cuda.synchronize()
    second = time.perf_counter() - t0
    audio_s = float(soffs_all[-1]) / SR
    out["config3_4096_mixed_1gpu"] = {"audio_seconds": audio_s, "tokens": int(total_tok), "sub_batches": len(subs), "wall_s_first": first, "wall_s": second,
                                      "audio_s_per_s": audio_s / second, "tflops_algorithmic": 0.798e15 / second / 1e12,
                                      "finite": bool(torch.isfinite(emb_all[:: 997]).all().item()),
                                      "note": "eager launches (every sub-batch has a different shape), host-side repacking and reorder copies included"}
    del audio, emb_all

    # ---- config 5: mel-frontend-only sweep vs HBM roofline (115 200 algorithmic bytes per audio-second)
    sweep = []
    for dur in (1, 10, 30, 60, 300, 1200):
        for batch in (1, 8, 64, 256, 1024):
            n = dur * SR
            if n * batch > (1 << 28):  # keep the device-resident input <= 1 GiB
                continue
            waves = [(0.1 * torch.randn(n, device="cuda"))] * batch
            ms = timed(lambda: log_mel_spectrogram_batch(waves), 10)
            gbs = 115200.0 * dur * batch / (ms / 1e3) / 1e9
            sweep.append({"duration_s": dur, "batch": batch, "ms": ms, "audio_s_per_s": dur * batch / (ms / 1e3), "gbs_algorithmic": gbs, "frac_hbm_peak": gbs / peaks["hbm_gbs"]})
    out["config5_mel_sweep"] = sweep
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "configs_r01.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps({k: v for k, v in out.items() if k != "config5_mel_sweep"}, indent=1))
    best = max(sweep, key=lambda r: r["gbs_algorithmic"])
    print("mel sweep best:", best)


if __name__ == "__main__":
    main()
