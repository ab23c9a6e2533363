"""Decoder-prefill throughput at the 1.7B architecture on one B200 (not a pytest file).

    python tests/run_prefill.py [--batch 64] [--audio-seconds 30] [--iters 5] [--out gpurun_out/prefill.json]

Workload: the prompts of BASELINE config 2 (64 x 30 s): 390 audio tokens + 17 prompt tokens (tokenizer.py:56-86 with a
one-token language name) = 407 rows per prompt, 26 048 rows per batch; random-init weights (seed 4321), synthetic embeddings.
Algorithmic FLOPs per prompt row and layer: 2 * (hidden * (q + 2 kv) + q * hidden + 3 * hidden * intermediate) = 100.7 M,
x 28 layers = 2.82 G, plus causal attention 2 * 2 * T^2 / 2 * q per layer per prompt, plus the last-token lm_head.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from qwen3_asr_mlx_b200 import decoder as dec  # noqa: E402
from qwen3_asr_mlx_b200.config import TextDecoderConfig  # noqa: E402
from qwen3_asr_mlx_b200.launcher import tokens_for_samples  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--audio-seconds", type=float, default=30.0)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "prefill.json"))
    args = ap.parse_args()
    cfg = TextDecoderConfig()
    d = dec.TextDecoder(cfg, device=0)
    d.load_weights(dec.random_init(cfg, seed=4321, device="cuda:0"))
    T = tokens_for_samples(int(args.audio_seconds * 16000)) + 17
    offs = np.arange(args.batch + 1, dtype=np.int64) * T
    n = int(offs[-1])
    emb = (torch.randn(n, cfg.hidden_size, device="cuda") * 0.05).bfloat16()
    for _ in range(2):
        d.prefill(emb, offs)
    torch.cuda.synchronize()
    l0 = d.stats()["kernel_launches"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        last, cache = d.prefill(emb, offs)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    H, Q, KV, I, L = cfg.hidden_size, cfg.num_attention_heads * cfg.head_dim, cfg.num_key_value_heads * cfg.head_dim, cfg.intermediate_size, cfg.num_hidden_layers
    gemm = 2.0 * n * L * (H * (Q + 2 * KV) + Q * H + 3 * H * I) + 2.0 * args.batch * H * cfg.vocab_size
    attn = L * args.batch * 2.0 * 2.0 * (T * (T + 1) / 2) * Q
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1365.6)) or 1365.6)
    out = {"workload": f"{args.batch} prompts x {T} rows (={n}), Qwen3-ASR-1.7B text decoder, random init, bf16 (fp32 accumulate)",
           "ms_per_prefill": ms, "prompt_rows_per_s": n / (ms / 1e3), "audio_s_per_s": args.batch * args.audio_seconds / (ms / 1e3),
           "gemm_tflop": gemm / 1e12, "attention_tflop": attn / 1e12, "tflops_algorithmic": (gemm + attn) / (ms / 1e3) / 1e12,
           "frac_of_bf16_peak": (gemm + attn) / (ms / 1e3) / 1e12 / peak, "peak_tflops": peak,
           "launches_per_prefill": (d.stats()["kernel_launches"] - l0) // args.iters,
           "kv_cache_bytes": int(cache.keys.numel() * 2 * 2), "finite": bool(torch.isfinite(last.tensor).all().item())}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(out, open(args.out, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
