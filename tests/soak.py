"""Soak run (not a pytest file): thousands of replayed calls of three shapes, results compared bit for bit with the first pass
every few hundred iterations; device memory must not grow with the iteration count (a constant ~57 MB appears once, when torch
loads the modules of the comparison kernels used inside the loop: 1000 and 5000 iterations show the same delta).  Catches rare ordering bugs (programmatic dependent launch, graph
replay, the double-buffered host pipeline) that a single test pass cannot.

    python tests/soak.py [iterations]
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import synth  # noqa: E402
from qwen3_asr_mlx_b200 import AudioEncoder, AudioEncoderConfig, weights  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
cfg = AudioEncoderConfig()
enc = AudioEncoder(cfg, device=0)
enc.load_weights(weights.random_init(cfg, seed=1234))
rng = np.random.default_rng(77)
shapes = [[16000 * 10], [16000 * 30] * 8, [int(n) for n in rng.integers(16000, 480001, size=24)]]
cases = []
for lens in shapes:
    x = torch.from_numpy(np.concatenate([synth(rng, n) for n in lens])).cuda()
    so = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=so[1:])
    ntok = sum(enc.num_tokens(n // 160) for n in lens)
    out = torch.empty((ntok, cfg.output_dim), dtype=torch.float32, device="cuda")
    for _ in range(3):
        enc.encode_packed_audio(x, so, out=out)
    cases.append((x, so, out, out.clone()))
# warm the host pipeline too (its two slots own device staging buffers), then take the memory baseline
x, so, out, want = cases[1]
pin_in = [x.cpu().pin_memory().numpy() for _ in range(2)]
pin_out = [torch.empty(tuple(out.shape), dtype=torch.float32).pin_memory().numpy() for _ in range(2)]
for s in (0, 1, 0, 1, 0, 1):  # eager, captured, replayed per slot: every graph exists before the baseline is taken
    enc.host_wait(s)
    enc.encode_audio_host_async(s, pin_in[s], so, pin_out[s])
enc.host_wait(0)
enc.host_wait(1)
torch.cuda.synchronize()
free0 = torch.cuda.mem_get_info()[0]
t0 = time.time()
bad = 0
for i in range(iters):
    x, so, out, want = cases[i % 3]
    enc.encode_packed_audio(x, so, out=out)
    if i % 97 == 0:
        bad += int(not torch.equal(out, want))
        out.zero_()
torch.cuda.synchronize()
free_mid = torch.cuda.mem_get_info()[0]
# host pipeline: 300 double-buffered steps of the 8 x 30 s case
x, so, out, want = cases[1]
want_np = want.cpu().numpy()
for i in range(300):
    s = i & 1
    enc.host_wait(s)
    if i >= 2 and i % 31 == 0:
        bad += int(not np.array_equal(pin_out[s], want_np))
    enc.encode_audio_host_async(s, pin_in[s], so, pin_out[s])
enc.host_wait(0)
enc.host_wait(1)
bad += int(not np.array_equal(pin_out[0], want_np)) + int(not np.array_equal(pin_out[1], want_np))
free1 = torch.cuda.mem_get_info()[0]
print(f"soak: {iters} replayed calls + 300 pipelined host steps in {time.time() - t0:.1f} s, mismatches {bad}, "
      f"device memory delta {(free0 - free_mid) / 1e6:.1f} MB after the replays, {(free0 - free1) / 1e6:.1f} MB at the end, launches {enc.stats()['kernel_launches']}")
sys.exit(1 if bad or free0 - free1 > 128e6 else 0)
