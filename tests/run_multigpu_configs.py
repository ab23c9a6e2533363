"""BASELINE.json configs[2] and configs[3] on N GPUs of one box (run under torchrun; not a pytest file):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 \
        tests/run_multigpu_configs.py [--out gpurun_out/multigpu_configs.json]

  config 3: 4096 mixed-length utterances (1-30 s, length seed 20261018 of SURVEY.md 8d), LPT-partitioned over the
            ranks by token count, varlen-packed sub-batches of <= 32768 tokens, no collective on the forward path,
            ONE final all-gather-v of the bf16 embeddings (restoring the original utterance order).
  config 4: one 20-minute file, cut at low-energy boundaries by the reference's splitter rule
            (`_find_split_points(samples, 30 s, +-5 s)`, model.py:400-403,454-513) into ~40 segments that are
            encoded data-parallel (per-segment mel max, like model.py:418) and gathered; plus the same file as a
            single pass (150 windows) on rank 0 for comparison.
Times are device-side (CUDA events) from a barrier to the end of the gather, max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from qwen3_asr_mlx_b200 import AudioEncoder, AudioEncoderConfig, launcher, weights  # noqa: E402
from qwen3_asr_mlx_b200.model import _find_split_points  # noqa: E402

SR = 16000


def timed_max(fn, world):
    """Device time of fn() between two barriers, max over ranks (ms)."""
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    result = fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), result


def long_file(seed=4, seconds=1200):
    """noise + tones with 0.5 s near-silent (x1e-3) gaps every ~7-13 s."""
    rng = np.random.default_rng(seed)
    n = seconds * SR
    t = np.arange(n, dtype=np.float32) / SR
    x = 0.1 * rng.standard_normal(n).astype(np.float32)
    for _ in range(3):
        x += (0.3 * np.sin(2 * np.pi * rng.uniform(100, 4000) * t + rng.uniform(0, 6.28))).astype(np.float32)
    pos = 0
    while pos < n:
        pos += int(rng.uniform(7.0, 13.0) * SR)
        x[pos: pos + SR // 2] *= 1e-3
    return np.clip(x, -1, 1).astype(np.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "multigpu_configs.json"))
    ap.add_argument("--utterances", type=int, default=4096)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = AudioEncoderConfig()
    enc = AudioEncoder(cfg, device=local)
    enc.load_weights(weights.random_init(cfg, seed=1234))
    out = {"world_size": world}

    # ------------------------------------------------------------------ config 3
    lengths = [int(n) for n in np.random.default_rng(20261018).integers(16000, 480001, size=args.utterances)]
    costs = [launcher.tokens_for_samples(n) for n in lengths]
    parts = launcher.lpt_partition(costs, world)
    mine = parts[rank]
    gen = torch.Generator(device="cuda").manual_seed(3)
    audio = {i: 0.1 * torch.randn(lengths[i], device="cuda", generator=gen) for i in mine}  # this rank's share only

    def encode_fn(idx):
        so = np.zeros(len(idx) + 1, dtype=np.int64)
        np.cumsum([lengths[i] for i in idx], out=so[1:])
        packed = torch.cat([audio[i] for i in idx])
        emb, toffs = enc.encode_packed_audio(packed, so, out_dtype="bfloat16")
        return emb.tensor, toffs

    def run3(gather):
        return launcher.encode_sharded(encode_fn, lengths, cfg.output_dim, rank, world, tokens_per_call=32768, gather=gather,
                                       dtype=torch.bfloat16)

    run3(False)  # warm-up: workspaces, first-touch
    ms_fwd, _ = timed_max(lambda: run3(False), world)
    ms_all, (emb, offs, _) = timed_max(lambda: run3(True), world)
    ms_p2p, same = None, None
    if world > 1:  # the same gather as ONE scatter kernel over NVLink peer memory (launcher.PeerGather)
        pg = launcher.PeerGather(int(offs[-1]), cfg.output_dim, dtype=torch.bfloat16)

        def run3_p2p():
            return launcher.encode_sharded(encode_fn, lengths, cfg.output_dim, rank, world, tokens_per_call=32768, gather=True,
                                           dtype=torch.bfloat16, peer_gather=pg)

        run3_p2p()
        ms_p2p, (emb_p, _, _) = timed_max(run3_p2p, world)
        same = bool(torch.equal(emb_p, emb))
    audio_s = sum(lengths) / SR
    per_rank_tokens = [sum(costs[i] for i in p) for p in parts]
    out["config3_mixed_length"] = {
        "utterances": len(lengths), "audio_seconds": audio_s, "tokens": int(offs[-1]), "tokens_per_rank_min_max": [min(per_rank_tokens), max(per_rank_tokens)],
        "ms_forward_only": ms_fwd, "audio_s_per_s_forward_only": audio_s / (ms_fwd / 1e3),
        "ms_with_final_gather": ms_all, "audio_s_per_s_with_final_gather": audio_s / (ms_all / 1e3),
        "ms_with_peer_memory_gather": ms_p2p, "audio_s_per_s_with_peer_memory_gather": (audio_s / (ms_p2p / 1e3)) if ms_p2p else None,
        "peer_memory_gather_equals_nccl_gather": same,
        "gathered_shape": list(emb.shape), "gathered_finite": bool(torch.isfinite(emb[::997].float()).all().item()),
        "gather_bytes_bf16": int(offs[-1]) * cfg.output_dim * 2,
        "note": "eager launches (every sub-batch has its own shape); host-side packing (torch.cat) inside the timed region",
    }
    del audio, emb

    # ------------------------------------------------------------------ config 4
    x = long_file()
    cuts = _find_split_points(x, 30 * SR, 5 * SR)
    bounds = [0] + cuts + [len(x)]
    segs = [(a, b) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]
    seg_len = [b - a for a, b in segs]
    xd = torch.from_numpy(x).cuda()

    def encode_seg(idx):
        so = np.zeros(len(idx) + 1, dtype=np.int64)
        np.cumsum([seg_len[i] for i in idx], out=so[1:])
        packed = torch.cat([xd[segs[i][0]: segs[i][1]] for i in idx])
        emb, toffs = enc.encode_packed_audio(packed, so, out_dtype="bfloat16")
        return emb.tensor, toffs

    def run4():
        return launcher.encode_sharded(encode_seg, seg_len, cfg.output_dim, rank, world, tokens_per_call=32768, gather=True, dtype=torch.bfloat16)

    run4()
    ms4, (emb4, offs4, _) = timed_max(run4, world)
    out["config4_20min_chunked"] = {"segments": len(segs), "segment_seconds_min_max": [min(seg_len) / SR, max(seg_len) / SR], "tokens": int(offs4[-1]),
                                    "ms": ms4, "audio_s_per_s": 1200.0 / (ms4 / 1e3), "finite": bool(torch.isfinite(emb4.float()).all().item())}
    # the default chunk_duration (1200 s) path: ONE utterance, single pass, its 150 attention windows sharded over the ranks
    if world > 1:
        from qwen3_asr_mlx_b200 import log_mel_spectrogram

        pg4 = launcher.PeerGather(15600, cfg.output_dim, dtype=torch.bfloat16)

        def run4_single_pass():
            mel = log_mel_spectrogram(xd).tensor  # every rank: full mel with the utterance-wide max (audio.py:275)
            return launcher.encode_long_sharded(enc, mel, rank, world, peer_gather=pg4, out_dtype="bfloat16")

        for _ in range(2):
            run4_single_pass()
        ms4s, emb4s = timed_max(run4_single_pass, world)
        out["config4_20min_single_pass_sharded"] = {"ms": ms4s, "audio_s_per_s": 1200.0 / (ms4s / 1e3), "tokens": int(emb4s.shape[0]),
                                                    "windows_per_rank": [-(-(b - a) // 800) for a, b in launcher.window_shares(120000, world)],
                                                    "note": "mel on every rank + window share + peer-memory gather, eager launches"}
    if rank == 0:  # the same on one GPU
        so = np.array([0, len(x)], dtype=np.int64)
        o = torch.empty((15600, cfg.output_dim), dtype=torch.bfloat16, device="cuda")
        for _ in range(3):
            enc.encode_packed_audio(xd, so, out_dtype="bfloat16", out=o)
        ms1, _ = timed_max(lambda: enc.encode_packed_audio(xd, so, out_dtype="bfloat16", out=o), 1)
        out["config4_20min_single_pass_rank0"] = {"ms": ms1, "audio_s_per_s": 1200.0 / (ms1 / 1e3), "tokens": 15600}
    if world > 1:
        dist.barrier()
    if rank == 0:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(out, f, indent=1)
        print(json.dumps(out, indent=1))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
