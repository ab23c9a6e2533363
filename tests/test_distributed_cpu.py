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
