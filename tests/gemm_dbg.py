import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["QASR_GEMM_DBG"] = "1"
from qwen3_asr_mlx_b200 import _lib
lib = _lib.load()
names = {0: "store_bf16", 1: "gelu_bf16", 2: "resid_f32", 3: "discard", 4: "math_only"}
for (M, N, K) in [(24960, 3072, 1024), (24960, 1024, 4096)]:
    for epi in (3, 4, 0, 1, 2):
        ms = ctypes.c_float()
        print(f"--- {M}x{N}x{K} pair {names[epi]}", file=sys.stderr, flush=True)
        _lib.check(lib.qasr_bench_gemm(0, M, N, K, epi + 16, 5, ctypes.byref(ms)))
        print(f"    {2.0 * M * N * K / (ms.value * 1e-3) / 1e12:7.1f} TF/s  {ms.value*1e3:.1f} us", file=sys.stderr, flush=True)
