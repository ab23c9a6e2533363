"""CPU, world_size 2 over gloo: the launcher's shard -> encode -> gather -> reorder path."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from qwen3_asr_mlx_b200 import launcher

N_SAMPLES = [16000 * k + 37 * k for k in (3, 1, 7, 2, 5, 30, 11, 4, 9)]
DIM = 8


def _expected():
    rows = []
    for i, n in enumerate(N_SAMPLES):
        t = launcher.tokens_for_samples(n)
        rows.append(torch.arange(t, dtype=torch.float32)[:, None] + 1000.0 * i + torch.zeros(DIM))
    return torch.cat(rows)


class _FakeEncoder:
    """encode_batch([mel]) -> one row per token, holding the first frame value of the token's chunk and the token's index
    within the chunk (tokens follow the reference's count rule)."""

    class config:
        output_dim = DIM

    def encode_batch(self, mels, out_dtype="float32"):
        from qwen3_asr_mlx_b200._array import DeviceArray

        mel = mels[0]
        rows = []
        for f0 in range(0, mel.shape[1], 100):
            n = launcher.conv_output_length(min(100, mel.shape[1] - f0))
            rows.append(torch.stack([mel[0, f0] + torch.zeros(DIM) + 0.01 * t for t in range(n)]))
        return DeviceArray(torch.cat(rows)), None


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        def fake_encode(idx):
            rows = []
            for i in idx:
                t = launcher.tokens_for_samples(N_SAMPLES[i])
                rows.append(torch.arange(t, dtype=torch.float32)[:, None] + 1000.0 * i + torch.zeros(DIM))
            return torch.cat(rows), np.cumsum([0] + [r.shape[0] for r in rows])

        emb, offs, mine = launcher.encode_sharded(fake_encode, N_SAMPLES, DIM, rank, world, tokens_per_call=200)
        ok = bool(torch.equal(emb, _expected())) and int(offs[-1]) == emb.shape[0]
        # one long utterance sharded by attention windows (config 4, single pass): a fake encoder whose token rows carry the
        # absolute frame of their chunk, so any mis-cut or mis-ordered share shows
        T = 800 * 5 + 333
        mel = torch.arange(T, dtype=torch.float32)[None, :].repeat(4, 1)
        long_emb = launcher.encode_long_sharded(_FakeEncoder(), mel, rank, world)
        ok = ok and bool(torch.equal(long_emb, _FakeEncoder().encode_batch([mel])[0].tensor))
        q.put((rank, ok, mine))
    finally:
        dist.destroy_process_group()


def test_two_rank_gather_restores_original_order():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in results)
    shares = {r: mine for r, _, mine in results}
    assert sorted(shares[0] + shares[1]) == list(range(len(N_SAMPLES))) and shares[0] and shares[1]
