"""tcgen05 GEMM micro-benchmark (not a pytest file): python tests/gemm_microbench.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qwen3_asr_mlx_b200 import _lib
lib = _lib.load()
names = {0: "store_bf16", 1: "gelu_bf16", 2: "resid_f32", 3: "discard", 4: "math_only"}
shapes = [(8192, 8192, 8192), (24960, 3072, 1024), (24960, 4096, 1024), (24960, 1024, 4096), (24960, 1024, 1024)]
for (M, N, K) in shapes:
    for pair in (0, 16):
        row = []
        for epi in ((3, 0, 1, 2) if pair == 0 else (3, 4, 0, 1, 2)):
            ms = ctypes.c_float()
            _lib.check(lib.qasr_bench_gemm(0, M, N, K, epi + pair, 20, ctypes.byref(ms)))
            row.append(f"{names[epi]} {2.0 * M * N * K / (ms.value * 1e-3) / 1e12:7.1f} TF/s")
        print(f"{M}x{N}x{K} cta_group::{2 if pair else 1}: " + " | ".join(row), flush=True)
