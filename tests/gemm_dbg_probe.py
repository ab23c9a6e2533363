"""Where does the MMA-issuing thread of the residual (fp32 TMA reduce-add) GEMM spend its time?  (not a pytest file)
    QASR_GEMM_DBG=1 python tests/gemm_dbg_probe.py
Prints, per epilogue, the throughput and the instrumented cycle accounting of CTA 0 / 74 (compile-time instrumented build)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qwen3_asr_mlx_b200 import _lib
lib = _lib.load()
names = {0: "store_bf16", 1: "gelu_bf16", 2: "resid_f32", 3: "discard", 4: "math_only"}
for (M, N, K) in [(24960, 1024, 1024), (24960, 1024, 4096), (24960, 3072, 1024)]:
    for epi in (3, 0, 2):
        ms = ctypes.c_float()
        sys.stderr.flush()
        _lib.check(lib.qasr_bench_gemm(0, M, N, K, epi + 16, 20, ctypes.byref(ms)))
        print(f"{M}x{N}x{K} {names[epi]}: {ms.value * 1e3:.1f} us  {2.0 * M * N * K / (ms.value * 1e-3) / 1e12:7.1f} TF/s", flush=True)
