"""Minimal driver for ncu captures of the decoder-prefill kernels: a 2-layer decoder with the 1.7B widths, 64 x 407 prompt rows
(BASELINE config 2's prompts), two prefill calls (ncu: `-s 21 -c 21` captures the second).  One call = 21 launches:
cast, 2 x [rmsnorm, qkv GEMM, qknorm_rope, causal attention, o_proj GEMM, rmsnorm, gate|up SwiGLU GEMM, down GEMM],
final rmsnorm (last rows), lm_head GEMM, + the hidden copy is a memcpy."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from qwen3_asr_mlx_b200 import decoder as dec  # noqa: E402
from qwen3_asr_mlx_b200.config import TextDecoderConfig  # noqa: E402

cfg = TextDecoderConfig(num_hidden_layers=2)
d = dec.TextDecoder(cfg, device=0)
d.load_weights(dec.random_init(cfg, seed=4321, device="cuda:0"))
offs = np.arange(65, dtype=np.int64) * 407
emb = (torch.randn(int(offs[-1]), cfg.hidden_size, device="cuda") * 0.05).bfloat16()
l0 = d.stats()["kernel_launches"]
for _ in range(2):
    last, cache = d.prefill(emb, offs)
torch.cuda.synchronize()
print("launches per call:", (d.stats()["kernel_launches"] - l0) // 2, "checksum", float(last.tensor[0, 0]))
