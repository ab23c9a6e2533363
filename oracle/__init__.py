"""CPU oracle of the Qwen3-ASR audio-encoding hot path.

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package.  The product (qwen3_asr_mlx_b200) never does.

Pinning status
* mel  : PINNED.  oracle/mel_np.py is checked bit-for-bit against the reference's own
         log_mel_spectrogram executed verbatim (oracle/mel_ref.py, fixtures in tests/golden/).
* encoder: PARITY UNPINNED by the reference.  MLX is not installable here, and the reference's
         tests hold shapes only (tests/test_encoder.py), no numeric vectors.  Two independent
         restatements (numpy fp64 loops, torch fp32 library ops) are cross-checked instead.
"""
