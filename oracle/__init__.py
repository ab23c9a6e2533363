"""CPU oracle of the Qwen3-ASR audio-encoding hot path.

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package.  The product (qwen3_asr_mlx_b200) never does.

Pinning status
* mel  : PINNED.  oracle/mel_np.py is checked bit-for-bit against the reference's own
         log_mel_spectrogram executed verbatim (oracle/mel_ref.py, fixtures in tests/golden/).
* encoder, decoder prefill, prompt assembly : PINNED to the reference's own code.  MLX itself is not installable
         here, so /root/reference/src/qwen3_asr_mlx is imported UNMODIFIED on top of a torch-CPU fp32 stand-in for
         the MLX calls it makes (oracle/_mlx_shim, driven by oracle/reference_ref.py); 97 of the reference's own
         tests (everything that does not need the downloaded checkpoint) pass on that stand-in.  Its outputs are
         committed as tests/golden/{encoder,decoder,prompt}_reference.npz (oracle/gen_golden.py) and the
         restatements here (encoder_torch.py, encoder_np.py, decoder_torch.py, prompt_np.py) equal them to <= 1e-5
         (measured 5e-7).  What remains assumed is the stand-in's reading of MLX's documented op semantics (listed in
         its headers); those are cross-checked against the model authors' PyTorch implementation in transformers
         (tests/test_oracle_upstream.py, tests/test_oracle_decoder.py).
"""
