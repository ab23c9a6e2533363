"""Per-kernel-category device time of small calls (not a pytest file): python tests/small_call_profile.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qwen3_asr_mlx_b200 import AudioEncoder, AudioEncoderConfig, weights
cfg = AudioEncoderConfig()
enc = AudioEncoder(cfg, device=0)
enc.load_weights(weights.random_init(cfg, seed=1234))
SR = 16000
for batch, seconds in ((1, 10), (1, 60), (8, 30)):
    n = seconds * SR
    x = 0.1 * torch.randn(batch * n, device="cuda")
    so = np.arange(batch + 1, dtype=np.int64) * n
    ntok = batch * enc.num_tokens(n // 160)
    out = torch.empty((ntok, cfg.output_dim), dtype=torch.float32, device="cuda")
    for _ in range(3):
        enc.encode_packed_audio(x, so, out=out)
    torch.cuda.synchronize()
    enc.set_profile(True)
    reps = 20
    for _ in range(reps):
        enc.encode_packed_audio(x, so, out=out)
    torch.cuda.synchronize()
    prof = enc.get_profile()
    enc.set_profile(False)
    tot = sum(v["ms"] for v in prof.values()) / reps
    print(f"{batch} x {seconds} s ({ntok} tokens): sum of kernels {tot:.3f} ms")
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        if v["launches"]:
            print(f"   {k:18s} {v['ms'] / reps:7.4f} ms  x{v['launches'] // reps:3d}  {1e3 * v['ms'] / v['launches']:6.2f} us/launch")
