"""The drop-in boundary is the C ABI alone: a plain-C11 program (tests/c_client/qasr_client.c, no Python / torch / C++ types)
compiled with gcc against include/qasr.h drives libqasr.so through the call sequence of the reference's call site
(model.py:331-335) and must produce the bits the Python host produces."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_client", "qasr_client.c")
LIBDIR = os.path.join(ROOT, "qwen3_asr_mlx_b200", "lib")
CFG_FIELDS = ("d_model", "encoder_layers", "encoder_attention_heads", "encoder_ffn_dim", "num_mel_bins", "max_source_positions",
              "output_dim", "n_window", "n_window_infer", "downsample_hidden_size")


def _build(tmp_path):
    exe = str(tmp_path / "qasr_client")
    cmd = ["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), SRC,
           "-L", LIBDIR, "-lqasr", f"-Wl,-rpath,{LIBDIR}", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def _write_inputs(tmp_path, cfg, params, audios):
    wpath, apath = str(tmp_path / "weights.bin"), str(tmp_path / "audio.f32")
    with open(wpath, "wb") as f:
        f.write(struct.pack("<10i", *[int(getattr(cfg, n)) for n in CFG_FIELDS]))
        f.write(struct.pack("<i", len(params)))
        for name, w in params.items():
            w = np.ascontiguousarray(w, dtype=np.float32)
            f.write(struct.pack("<i", len(name)) + name.encode() + struct.pack("<i", w.ndim) + struct.pack(f"<{w.ndim}q", *w.shape))
            f.write(w.tobytes())
    soffs = np.zeros(len(audios) + 1, dtype=np.int64)
    np.cumsum([len(a) for a in audios], out=soffs[1:])
    with open(apath, "wb") as f:
        f.write(struct.pack("<i", len(audios)) + soffs.tobytes() + np.concatenate(audios).astype(np.float32).tobytes())
    return wpath, apath


def test_header_is_plain_c_and_the_client_links(tmp_path):
    """include/qasr.h compiles as pedantic C11, every symbol the client uses resolves in libqasr.so, and without a GPU the
    library fails loudly through its error convention (no CPU fallback)."""
    import torch

    from qwen3_asr_mlx_b200 import AudioEncoderConfig, weights

    exe = _build(tmp_path)
    assert subprocess.run([exe], capture_output=True).returncode == 2  # usage
    if torch.cuda.is_available():
        return
    cfg = AudioEncoderConfig(d_model=128, encoder_layers=1, encoder_attention_heads=2, encoder_ffn_dim=256, output_dim=128)
    wpath, apath = _write_inputs(tmp_path, cfg, weights.random_init(cfg, seed=1), [np.zeros(1600, dtype=np.float32)])
    r = subprocess.run([exe, wpath, apath, str(tmp_path / "out.f32")], capture_output=True, text=True)
    assert r.returncode == 1 and "qasr_create" in r.stderr, (r.returncode, r.stderr)


@pytest.mark.gpu
def test_c_client_matches_the_python_host_bit_for_bit(tmp_path):
    from helpers import synth
    from qwen3_asr_mlx_b200 import AudioEncoder, AudioEncoderConfig, weights

    cfg = AudioEncoderConfig(d_model=256, encoder_layers=2, encoder_attention_heads=4, encoder_ffn_dim=512, output_dim=256)
    params = weights.random_init(cfg, seed=11, exercise_all=True)
    rng = np.random.default_rng(5)
    audios = [synth(rng, n) for n in (16000 * 7 + 123, 3000, 16000 * 21)]
    exe = _build(tmp_path)
    wpath, apath = _write_inputs(tmp_path, cfg, params, audios)
    opath = str(tmp_path / "out.f32")
    r = subprocess.run([exe, wpath, apath, opath], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "qasr_client ok" in r.stdout
    raw = open(opath, "rb").read()
    toffs = np.frombuffer(raw[: 8 * (len(audios) + 1)], dtype=np.int64)
    emb = np.frombuffer(raw[8 * (len(audios) + 1):], dtype=np.float32).reshape(-1, cfg.output_dim)
    enc = AudioEncoder(cfg)
    enc.load_weights(params)
    want, want_offs = enc.encode_audio_batch(audios)
    assert np.array_equal(toffs, want_offs)
    assert np.array_equal(emb.view(np.uint32), np.array(want).view(np.uint32))
    enc.close()
