"""Graph-replayed mel + encoder time for a few call shapes (not a pytest file): QASR_PDL=0/1 python tests/pdl_probe.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qwen3_asr_mlx_b200 import AudioEncoder, AudioEncoderConfig, weights
cfg = AudioEncoderConfig()
enc = AudioEncoder(cfg, device=0)
enc.load_weights(weights.random_init(cfg, seed=1234))
SR = 16000
res = []
for batch, seconds in ((1, 10), (1, 60), (1, 300), (8, 30), (1, 1200), (32, 30), (64, 30)):
    n = seconds * SR
    x = 0.1 * torch.randn(batch * n, device="cuda")
    so = np.arange(batch + 1, dtype=np.int64) * n
    ntok = batch * enc.num_tokens(n // 160)
    out = torch.empty((ntok, cfg.output_dim), dtype=torch.bfloat16, device="cuda")
    for _ in range(4):
        enc.encode_packed_audio(x, so, out_dtype="bfloat16", out=out)
    torch.cuda.synchronize()
    reps = 200 if ntok < 2000 else 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        for _ in range(reps):
            enc.encode_packed_audio(x, so, out_dtype="bfloat16", out=out)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    res.append(f"{batch}x{seconds}s({ntok} tok)={best:.3f}")
print("PDL=" + os.environ.get("QASR_PDL", "1"), " ".join(res))
