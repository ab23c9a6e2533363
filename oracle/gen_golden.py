"""Generate tests/golden/* (run in the authoring container, where /root/reference exists).

mel_*.npz     : inputs and outputs of the REFERENCE's own log_mel_spectrogram executed verbatim
                (oracle/mel_ref.py) -> these pin oracle/mel_np.py and the CUDA mel kernels.
mel_filterbank.npy : the reference's _get_mel_filterbank().
encoder_small.npz  : oracle/encoder_np.py (fp64) output for a small seeded configuration; a drift anchor for
                the two restatements.
encoder_reference.npz, decoder_reference.npz, prompt_reference.npz, language_map_reference.json :
                outputs of the REFERENCE's own encoder.py / decoder.py / generate.prepare_inputs /
                tokenizer.build_prompt / model.LANGUAGE_MAP executed unmodified behind the torch-backed MLX stand-in
                (oracle/reference_ref.py) -> these pin oracle/encoder_{torch,np}.py, decoder_torch.py, prompt_np.py
                and the CUDA path.  Inputs are regenerated from seeds by tests/reference_cases.py (shared with the
                tests); a float64 checksum of every input is stored beside the output.

    python oracle/gen_golden.py               # everything
    python oracle/gen_golden.py reference     # only the *_reference fixtures of encoder / decoder / prompt
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import encoder_np, mel_np, mel_ref  # noqa: E402
from qwen3_asr_mlx_b200 import weights  # noqa: E402
from qwen3_asr_mlx_b200.config import AudioEncoderConfig  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def synth(rng, n):
    """SURVEY.md §8d recipe: 0.1*N(0,1) + 3 tones (100-4000 Hz, amp 0.3) x slow envelope, clipped."""
    t = np.arange(n) / 16000.0
    x = 0.1 * rng.standard_normal(n)
    for _ in range(3):
        x += 0.3 * np.sin(2 * np.pi * rng.uniform(100, 4000) * t + rng.uniform(0, 6.28)) * (0.5 + 0.5 * np.sin(2 * np.pi * rng.uniform(0.1, 1.0) * t))
    return np.clip(x, -1, 1).astype(np.float32)


def gen_reference_fixtures():
    """Run the reference's encoder / decoder / prepare_inputs verbatim on the seeded cases of tests/reference_cases.py."""
    import json

    from oracle import reference_ref
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import reference_cases as rc

    assert reference_ref.available(), "needs /root/reference"
    out = {}
    for group, (cfg, params) in rc.encoder_groups().items():
        enc = reference_ref.build_encoder(params, cfg)
        for name, mel in rc.encoder_inputs(group).items():
            out[f"{group}/{name}/emb"] = reference_ref.encoder_forward(params, cfg, mel, enc)
            out[f"{group}/{name}/checksum"] = np.array(rc.checksum(mel))
            print("encoder", group, name, mel.shape, out[f"{group}/{name}/emb"].shape, flush=True)
        del enc, params
    np.savez_compressed(os.path.join(GOLDEN, "encoder_reference.npz"), **out)

    out = {}
    for group, (cfg, params) in rc.decoder_groups().items():
        dec = reference_ref.build_decoder(params, cfg)
        for name, emb in rc.decoder_inputs(group).items():
            r = reference_ref.decoder_prefill(params, cfg, emb, dec)
            for k in ("logits", "keys", "values"):
                out[f"{group}/{name}/{k}"] = r[k]
            out[f"{group}/{name}/checksum"] = np.array(rc.checksum(emb))
            print("decoder", group, name, emb.shape, r["logits"].shape, r["keys"].shape, flush=True)
    np.savez_compressed(os.path.join(GOLDEN, "decoder_reference.npz"), **out)

    out = {}
    pkg = reference_ref.package()
    for name, (audio, lang_tokens, table) in rc.prompt_cases().items():
        ids = pkg.build_prompt(audio.shape[1], lang_tokens)  # the reference's own tokenizer.build_prompt (tokenizer.py:56-86)
        out[f"{name}/ids"] = np.array(ids, dtype=np.int64)
        out[f"{name}/embeds"] = reference_ref.prepare_inputs(audio, ids, table)
        out[f"{name}/checksum"] = np.array(rc.checksum(audio) + rc.checksum(table))
    np.savez_compressed(os.path.join(GOLDEN, "prompt_reference.npz"), **out)
    with open(os.path.join(GOLDEN, "language_map_reference.json"), "w") as f:
        json.dump(dict(pkg.LANGUAGE_MAP), f, indent=0, sort_keys=True, ensure_ascii=False)


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    assert mel_ref.available(), "needs /root/reference"
    if len(sys.argv) > 1 and sys.argv[1] == "reference":
        gen_reference_fixtures()
        return
    rng = np.random.default_rng(20261018)
    cases = {}
    for n in (160, 161, 199, 200, 201, 319, 400, 2417, 16000):
        cases[f"noise_{n}"] = (0.1 * rng.standard_normal(n)).astype(np.float32)
    cases["silence_16000"] = np.zeros(16000, dtype=np.float32)
    t = np.linspace(0.0, 1.0, 16000, endpoint=False)
    cases["tone440_16000"] = np.sin(2.0 * np.pi * 440.0 * t).astype(np.float32)
    cases["synth_48000"] = synth(rng, 48000)
    cases["synth_quiet_24000"] = (1e-3 * synth(rng, 24000)).astype(np.float32)
    out = {}
    for name, x in cases.items():
        out["in_" + name] = x
        out["out_" + name] = np.asarray(mel_ref.log_mel_spectrogram(x), dtype=np.float32)
    np.savez_compressed(os.path.join(GOLDEN, "mel_reference.npz"), **out)
    np.save(os.path.join(GOLDEN, "mel_filterbank.npy"), np.asarray(mel_ref.mel_filterbank(), dtype=np.float32))

    cfg = AudioEncoderConfig(d_model=256, encoder_layers=2, encoder_attention_heads=4, encoder_ffn_dim=512, output_dim=256)
    P = weights.random_init(cfg, seed=7, exercise_all=True)
    x = synth(np.random.default_rng(7), 16000 * 9 + 4321)  # 927 frames: 9 full chunks + 27 frames -> 117 + 4 = 121 tokens, 2 windows
    mel = mel_np.log_mel_spectrogram(x)
    emb = encoder_np.encoder_forward(P, cfg, mel)
    np.savez_compressed(os.path.join(GOLDEN, "encoder_small.npz"), audio=x, emb=emb.astype(np.float32),
                        cfg=np.array([cfg.d_model, cfg.encoder_layers, cfg.encoder_attention_heads, cfg.encoder_ffn_dim, cfg.output_dim]),
                        seed=np.array(7))
    # _find_split_points (model.py:454-513) executed verbatim: the function is pure numpy, but its module imports
    # mlx at top level, so the def is extracted with ast and exec'd alone.
    import ast
    src = open("/root/reference/src/qwen3_asr_mlx/model.py").read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "_find_split_points")
    ns = {"np": np}
    exec(compile(ast.Module(body=[node], type_ignores=[]), "reference_model_py", "exec"), ns)
    ref_split = ns["_find_split_points"]
    cases = {}
    # inputs are regenerated from the seed by the test (numpy's PCG64 stream is stable), only the answers are stored
    for i, (name, (n, chunk, search)) in enumerate({"a": (16000 * 25, 160000, 80000), "b": (16000 * 61 + 123, 480000, 80000),
                                                   "c": (16000 * 7, 16000, 32000), "d": (16000 * 33, 160000, 16000),
                                                   "e": (5000, 1000, 480)}.items()):
        r = np.random.default_rng(4200 + i)
        x = (r.standard_normal(n) * np.abs(np.sin(np.arange(n) / 9000.0))).astype(np.float32)
        cases[name + "_args"] = np.array([n, chunk, search, 4200 + i])
        cases[name + "_points"] = np.array(ref_split(x, chunk, search), dtype=np.int64)
    np.savez_compressed(os.path.join(GOLDEN, "split_points_reference.npz"), **cases)
    # frame energies with the reference's own expression (model.py:489-495) for the device kernel's bit-exact check,
    # plus the reference's split points on a config-4-style file (noise/tones with near-silent gaps; 120 s, 30 s chunks)
    # and with a non-default frame size (generic summation path)
    en = {}
    for i, (name, (n, chunk, search, frame)) in enumerate({"p": (16000 * 120 + 77, 480000, 80000, 480), "q": (16000 * 40, 160000, 80000, 400),
                                                          "r": (16000 * 21 + 5, 100000, 30000, 1000), "s": (479, 100, 100, 480),
                                                          "t": (16000 * 9, 16000, 32000, 480)}.items()):
        r = np.random.default_rng(5200 + i)
        x = (0.1 * r.standard_normal(n)).astype(np.float32)
        pos = 0
        while pos < n:
            pos += int(r.uniform(3.0, 6.0) * 16000)
            x[pos: pos + 8000] *= np.float32(1e-3)
        nf = n // frame
        en[name + "_args"] = np.array([n, chunk, search, frame, 5200 + i])
        en[name + "_energy"] = np.array([np.sqrt(np.mean(x[j * frame: (j + 1) * frame] ** 2)) for j in range(nf)], dtype=np.float32)
        en[name + "_points"] = np.array(ref_split(x, chunk, search, frame), dtype=np.int64)
    np.savez_compressed(os.path.join(GOLDEN, "split_energy_reference.npz"), **en)
    # load_audio (audio.py:173-204) executed verbatim on WAV files the tests rebuild byte-for-byte (tests/helpers.py:make_wav)
    import tempfile
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import WAV_CASES, make_wav
    la = {}
    with tempfile.TemporaryDirectory() as td:
        for name, kw in WAV_CASES.items():
            path = os.path.join(td, name + ".wav")
            open(path, "wb").write(make_wav(**kw))
            la[name] = np.asarray(mel_ref.module().load_audio(path), dtype=np.float32)
    np.savez_compressed(os.path.join(GOLDEN, "load_audio_reference.npz"), **la)
    gen_reference_fixtures()
    for f in sorted(os.listdir(GOLDEN)):
        print(f, os.path.getsize(os.path.join(GOLDEN, f)))


if __name__ == "__main__":
    main()
