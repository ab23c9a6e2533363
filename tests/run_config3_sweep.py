"""Experiment (not a pytest file): config 3 on ONE GPU as a function of the sub-batch size (tokens per libqasr call).
    python tests/run_config3_sweep.py [--utterances 4096]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qwen3_asr_mlx_b200 import AudioEncoder, AudioEncoderConfig, launcher, weights  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--utterances", type=int, default=4096)
ap.add_argument("--budgets", default="16384,32768,65536,131072")
args = ap.parse_args()
cfg = AudioEncoderConfig()
enc = AudioEncoder(cfg)
enc.load_weights(weights.random_init(cfg, seed=1234))
lengths = [int(n) for n in np.random.default_rng(20261018).integers(16000, 480001, size=args.utterances)]
audio = 0.1 * torch.randn(sum(lengths), device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
out = {}
for budget in [int(b) for b in args.budgets.split(",")]:
    def run():
        return launcher.encode_contiguous_sharded(enc, audio, lengths, 0, 1, gather=None, tokens_per_call=budget)
    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    out[budget] = {"ms": ms, "audio_s_per_s": sum(lengths) / 16000 / (ms / 1e3)}
    print(budget, out[budget], flush=True)
print(json.dumps(out))
