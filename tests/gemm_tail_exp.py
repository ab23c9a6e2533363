import ctypes, os, sys
sys.path.insert(0, "/root/repo")
from qwen3_asr_mlx_b200 import _lib
lib = _lib.load()
for K in (1024, 4096):
    for M in (18944, 24960, 28416, 26112):
        for N in (1024,):
            ms = ctypes.c_float()
            _lib.check(lib.qasr_bench_gemm(0, M, N, K, 2 + 16, 200, ctypes.byref(ms)))
            tiles = ((M + 255) // 256) * (N // 256)
            print(f"M={M} N={N} K={K} tiles={tiles} waves={tiles/74:.2f} ms={ms.value:.4f} TF/s={2.0*M*N*K/(ms.value*1e-3)/1e12:7.1f} ms_per_wave_ceil={ms.value/ -(-tiles//74):.4f}", flush=True)
