"""Stand-in for the `mlx` package (TEST INFRASTRUCTURE ONLY, authoring container): torch-CPU fp32 implementations of
the `mlx.core` / `mlx.nn` calls the reference makes, so that /root/reference/src/qwen3_asr_mlx imports and runs
unmodified (oracle/mel_ref.py, oracle/reference_ref.py)."""
