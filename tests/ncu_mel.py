"""Minimal driver for an ncu capture of the mel kernel alone (config 2 batch): python tests/ncu_mel.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_workload
from qwen3_asr_mlx_b200 import log_mel_spectrogram_packed
audio, soffs = make_workload(0)
x = torch.from_numpy(audio).cuda()
for _ in range(3):
    m = log_mel_spectrogram_packed(x, soffs)
torch.cuda.synchronize()
print("ok")
