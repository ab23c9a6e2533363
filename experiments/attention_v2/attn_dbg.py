"""Experiment (not a pytest file): cycle accounting of window_attention_sm100_v2 from the instrumented library.
    QASR_LIB_PATH=qwen3_asr_mlx_b200/lib/libqasr_dbg.so QASR_GRAPHS=0 python tests/attn_dbg.py"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("QASR_GRAPHS", "0")
from bench import UTTS_PER_GPU, make_workload  # noqa: E402
from qwen3_asr_mlx_b200 import AudioEncoder, AudioEncoderConfig, _lib, weights  # noqa: E402

cfg = AudioEncoderConfig(encoder_layers=4)
enc = AudioEncoder(cfg, device=0)
enc.load_weights(weights.random_init(cfg, seed=1234))
audio, soffs = make_workload(0)
audio_dev = torch.from_numpy(audio).cuda()
out = torch.empty((UTTS_PER_GPU * 390, cfg.output_dim), dtype=torch.float32, device="cuda")
lib = _lib.load()
lib.qasr_debug_counters.restype = ctypes.c_int
buf = (ctypes.c_ulonglong * 32)()
for it in range(3):
    enc.encode_packed_audio(audio_dev, soffs, out=out)
    torch.cuda.synchronize()
    lib.qasr_debug_counters(buf, 32)
    c = list(buf)
    n = max(c[7], 1)
    names = ["wait s_full", "LDTM S", "max+exp+pack", "STTM+zero+arrive", "wait o_full", "LDTM O+arrive", "store"]
    print(f"iter {it}: group-0 thread items={c[7]}  cycles/item: " + ", ".join(f"{nm} {c[i] / n:.0f}" for i, nm in enumerate(names)),
          f"| total/item {sum(c[:7]) / n:.0f}")
    print(f"   QK thread: {c[8]} cycles for {c[9]} items ({c[8] / max(c[9], 1):.0f}/item); producer wait sempty {c[11] / max(c[9], 1):.0f}/item; "
          f"issue->complete latency: QK {c[13] / max(c[9], 1):.0f}, PV {c[12] / max(c[9], 1):.0f} cycles")
