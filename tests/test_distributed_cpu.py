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


class _FakePackedEncoder:
    """encode_packed_audio(audio, sample_offsets, out=...) writing one row per token: 1000 * (utterance id stored in the
    utterance's first sample) + token index, i.e. the rows of _expected()."""

    class config:
        output_dim = DIM

    def encode_packed_audio(self, audio, soffs, out_dtype="float32", out=None):
        row = 0
        for a, b in zip(soffs[:-1], soffs[1:]):
            t = launcher.tokens_for_samples(int(b - a))
            out[row: row + t] = torch.arange(t, dtype=torch.float32)[:, None] + 1000.0 * float(audio[int(a)]) + torch.zeros(DIM)
            row += t
        return out, np.concatenate([[0], np.cumsum([launcher.tokens_for_samples(int(b - a)) for a, b in zip(soffs[:-1], soffs[1:])])])

    # tail-block protocol of the launcher: the "hidden states" of the last call stay in the encoder, the projector fills
    # row blocks of the caller's choice
    def encode_packed_audio_hidden(self, audio, soffs):
        toffs = np.concatenate([[0], np.cumsum([launcher.tokens_for_samples(int(b - a)) for a, b in zip(soffs[:-1], soffs[1:])])])
        self._hidden = torch.empty((int(toffs[-1]), DIM))
        self.encode_packed_audio(audio, soffs, out=self._hidden)
        self.hidden_calls = getattr(self, "hidden_calls", 0) + 1
        return int(toffs[-1]), toffs

    def project_rows(self, row0, out):
        out.copy_(self._hidden[row0: row0 + out.shape[0]])
        self.projected_blocks = getattr(self, "projected_blocks", 0) + 1
        return out


class _GlooBlockGather:
    """Stand-in for launcher.PeerBlockGather on CPU: same begin / rows / push / finish protocol, the pushes are delivered
    with an all_gather_object in finish()."""

    def __init__(self, rows):
        self.buf = torch.full((rows, DIM), -1.0)
        self.dtype = torch.float32
        self.blocks = []
        self.peers = [r for r in range(dist.get_world_size()) if r != dist.get_rank()]

    def begin(self):
        dist.barrier()

    def rows(self, row0, n):
        return self.buf[row0: row0 + n]

    def push(self, row0, n):
        self.blocks.append((row0, self.buf[row0: row0 + n].clone()))

    def finish(self, total):
        everyone = [None] * dist.get_world_size()
        dist.all_gather_object(everyone, self.blocks)
        for blocks in everyone:
            for row0, t in blocks:
                self.buf[row0: row0 + t.shape[0]] = t
        self.blocks = []
        return self.buf[:total]


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
        # contiguous token-balanced shares + block pushes (the NVLink DMA gather of config 3): every rank packs ITS share only
        costs = [launcher.tokens_for_samples(n) for n in N_SAMPLES]
        share = launcher.contiguous_partition(costs, world)[rank]
        packed = torch.cat([torch.full((N_SAMPLES[i],), float(i)) for i in share])
        g = _GlooBlockGather(sum(costs))
        fake = _FakePackedEncoder()
        for _ in range(2):  # the buffer is reused by a second gather
            emb2, offs2, mine2 = launcher.encode_contiguous_sharded(fake, packed, N_SAMPLES, rank, world, gather=g,
                                                                    tokens_per_call=150, out_dtype="float32", tail_min_rows=32)
            ok = ok and mine2 == share and bool(torch.equal(emb2, _expected())) and int(offs2[-1]) == emb2.shape[0]
        # the last sub-batch of every pass went through the hidden-state call and was projected + pushed block by block
        ok = ok and getattr(fake, "hidden_calls", 0) == 2 and getattr(fake, "projected_blocks", 0) >= 4
        loc, _, _ = launcher.encode_contiguous_sharded(_FakePackedEncoder(), packed, N_SAMPLES, rank, world, gather=None, out_dtype="float32")
        ok = ok and bool(torch.equal(loc, _expected()[int(offs2[share[0]]): int(offs2[share[-1] + 1])]))
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
