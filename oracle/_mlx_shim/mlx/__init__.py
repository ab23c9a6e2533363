"""Two-line stand-in for the `mlx` package so that the reference's numpy-only audio.py imports.
TEST INFRASTRUCTURE ONLY (used by oracle/mel_ref.py inside the authoring container)."""
