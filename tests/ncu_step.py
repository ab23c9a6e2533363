"""Minimal driver for ncu captures: N eager steps of mel + encoder on BASELINE config 2 (64 x 30 s),
nothing else (no graphs, no profiling events), so that `-s <launches per step>` skips exactly one step.

    QASR_GRAPHS=0 python tests/ncu_step.py [steps]
One step = 181 kernel launches: 2 mel + 4 stem (conv1, conv2, conv3, conv_out; 2 chunk groups -> 8)
... printed at exit.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("QASR_GRAPHS", "0")

from bench import UTTS_PER_GPU, make_workload  # noqa: E402
from qwen3_asr_mlx_b200 import AudioEncoder, AudioEncoderConfig, weights  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg = AudioEncoderConfig()
enc = AudioEncoder(cfg, device=0)
enc.load_weights(weights.random_init(cfg, seed=1234))
audio, soffs = make_workload(0)
audio_dev = torch.from_numpy(audio).cuda()
out = torch.empty((UTTS_PER_GPU * 390, cfg.output_dim), dtype=torch.float32, device="cuda")
l0 = enc.stats()["kernel_launches"]
for _ in range(steps):
    enc.encode_packed_audio(audio_dev, soffs, out=out)
torch.cuda.synchronize()
print("launches per step:", (enc.stats()["kernel_launches"] - l0) // steps, "checksum", float(out[0, 0]))
