"""CPU: property tests (hypothesis) of the data-parallel launcher's host logic: LPT partition, sub-batch
budgeting, the token-count rule against the C library, and the row map that restores the original order after the gather."""
import ctypes

import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from qwen3_asr_mlx_b200 import _lib, launcher

costs_st = st.lists(st.integers(min_value=1, max_value=400), min_size=0, max_size=60)


@settings(max_examples=200, deadline=None)
@given(costs=costs_st, world=st.integers(min_value=1, max_value=8))
def test_lpt_partition_covers_everything_once_and_is_balanced(costs, world):
    parts = launcher.lpt_partition(costs, world)
    assert len(parts) == world
    flat = sorted(i for p in parts for i in p)
    assert flat == list(range(len(costs)))
    assert all(p == sorted(p) for p in parts)
    loads = [sum(costs[i] for i in p) for p in parts]
    if costs:
        # greedy LPT: no rank exceeds the lightest one by more than the heaviest single item
        assert max(loads) - min(loads) <= max(costs)
    assert parts == launcher.lpt_partition(costs, world)  # deterministic: every rank computes the same assignment


@settings(max_examples=300, deadline=None)
@given(costs=costs_st, world=st.integers(min_value=1, max_value=8))
def test_contiguous_partition_is_ordered_contiguous_and_balanced(costs, world):
    parts = launcher.contiguous_partition(costs, world)
    assert len(parts) == world
    assert [i for p in parts for i in p] == list(range(len(costs)))  # original order, every utterance once, contiguous runs
    if costs:
        share = sum(costs) / world
        for p in parts:  # a run's load differs from the ideal share by less than one (the heaviest) utterance
            assert abs(sum(costs[i] for i in p) - share) <= max(costs)
    assert parts == launcher.contiguous_partition(costs, world)


@settings(max_examples=200, deadline=None)
@given(costs=costs_st, budget=st.integers(min_value=1, max_value=1000))
def test_split_by_budget_keeps_order_and_budget(costs, budget):
    idx = list(range(len(costs)))
    subs = launcher.split_by_budget(idx, costs, budget)
    assert [i for s in subs for i in s] == idx
    for s in subs:
        assert s and (len(s) == 1 or sum(costs[i] for i in s) <= budget)


@settings(max_examples=300, deadline=None)
@given(n=st.integers(min_value=160, max_value=20_000_000))
def test_token_rule_matches_the_library(n):
    lib = _lib.load()
    frames, tokens = ctypes.c_int64(), ctypes.c_int64()
    assert lib.qasr_count_frames(n, ctypes.byref(frames)) == 0 and frames.value == n // 160
    assert lib.qasr_count_tokens(None, frames.value, ctypes.byref(tokens)) == 0
    assert tokens.value == launcher.tokens_for_samples(n)
    full, rem = divmod(n // 160, 100)
    assert tokens.value == 13 * full + (launcher.conv_output_length(rem) if rem else 0)


@settings(max_examples=200, deadline=None)
@given(costs=st.lists(st.integers(min_value=1, max_value=50), min_size=1, max_size=40), world=st.integers(min_value=1, max_value=6))
def test_gather_row_map_restores_original_order(costs, world):
    """Simulate the all-gathered buffer (rank r's rows at r * pad, its utterances back to back) and check that the row map
    puts every utterance's rows back at its original offset."""
    parts = launcher.lpt_partition(costs, world)
    c = np.asarray(costs, dtype=np.int64)
    per_rank = [int(c[p].sum()) if p else 0 for p in parts]
    pad = max(per_rank)
    flat = np.full(world * pad, -1, dtype=np.int64)  # each row holds (utterance id * 1000 + row within the utterance)
    for r, p in enumerate(parts):
        pos = r * pad
        for i in p:
            flat[pos: pos + costs[i]] = i * 1000 + np.arange(costs[i])
            pos += costs[i]
    start = launcher.gather_start_rows(parts, c, pad)
    offsets = np.concatenate([[0], np.cumsum(c)])
    src = np.repeat(start - offsets[:-1], c) + np.arange(int(c.sum()))
    restored = flat[src]
    want = np.concatenate([i * 1000 + np.arange(costs[i]) for i in range(len(costs))])
    assert np.array_equal(restored, want)


@settings(max_examples=200, deadline=None)
@given(costs=st.lists(st.integers(min_value=1, max_value=50), min_size=1, max_size=40), world=st.integers(min_value=1, max_value=8))
def test_peer_scatter_rows_tile_the_gathered_matrix(costs, world):
    """The destination rows of all ranks' local rows (PeerGather) are a permutation of 0..total-1, and every utterance lands
    at its original offset in order."""
    parts = launcher.lpt_partition(costs, world)
    offsets = np.concatenate([[0], np.cumsum(costs)])
    seen = np.full(int(offsets[-1]), -1, dtype=np.int64)
    for r, p in enumerate(parts):
        dst = launcher.scatter_dst_rows(p, costs)
        assert len(dst) == sum(costs[i] for i in p)
        payload = np.concatenate([i * 1000 + np.arange(costs[i]) for i in p]) if p else np.zeros(0, dtype=np.int64)
        assert (seen[dst] == -1).all()
        seen[dst] = payload
    want = np.concatenate([i * 1000 + np.arange(costs[i]) for i in range(len(costs))])
    assert np.array_equal(seen, want)


@settings(max_examples=200, deadline=None)
@given(T=st.integers(min_value=1, max_value=200_000), world=st.integers(min_value=1, max_value=8))
def test_window_shares_tile_the_utterance_on_the_window_grid(T, world):
    shares = launcher.window_shares(T, world)
    assert len(shares) == world and shares[0][0] == 0 and shares[-1][1] == T
    for (a, b), (c, d) in zip(shares[:-1], shares[1:]):
        assert b == c and a <= b
    assert all(a % 800 == 0 for a, _ in shares if a < T)
    # the token counts of the shares add up to the utterance's (no window is cut)
    assert sum(launcher.tokens_for_samples((b - a) * 160) for a, b in shares if b > a) == launcher.tokens_for_samples(T * 160)
    n_win = [-(-(b - a) // 800) for a, b in shares]
    assert max(n_win) - min(n_win) <= 1


class _RecordingGather:
    """PeerBlockGather protocol without any device: records the pushed row blocks."""

    def __init__(self, rows, dim, n_peers):
        import torch

        self.buf = torch.full((rows, dim), -1.0)
        self.dtype = torch.float32
        self.peers = list(range(n_peers))
        self.pushed = []

    def begin(self):
        self.pushed = []

    def rows(self, row0, n):
        return self.buf[row0: row0 + n]

    def push(self, row0, n):
        self.pushed.append((row0, n))

    def finish(self, total):
        return self.buf[:total]


class _RowIdEncoder:
    """Writes the GLOBAL token index of this rank's share into every output row (feature 0), through both the one-call path
    and the hidden-state + row-block projector path of the launcher."""

    class config:
        output_dim = 2

    def __init__(self):
        self.calls = []

    def _fill(self, soffs, out, start):
        import torch

        n = sum(launcher.tokens_for_samples(int(b - a)) for a, b in zip(soffs[:-1], soffs[1:]))
        out[:n, 0] = torch.arange(start, start + n, dtype=torch.float32)
        out[:n, 1] = 7.0
        return n

    def encode_packed_audio(self, audio, soffs, out_dtype="float32", out=None):
        n = self._fill(soffs, out, int(audio[0]))
        self.calls.append(("full", n))
        return out, np.array([0, n])

    def encode_packed_audio_hidden(self, audio, soffs):
        import torch

        n = sum(launcher.tokens_for_samples(int(b - a)) for a, b in zip(soffs[:-1], soffs[1:]))
        self._hidden = torch.zeros((n, 2))
        self._fill(soffs, self._hidden, int(audio[0]))
        self.calls.append(("hidden", n))
        return n, np.array([0, n])

    def project_rows(self, row0, out):
        out.copy_(self._hidden[row0: row0 + out.shape[0]])
        self.calls.append(("block", int(out.shape[0])))
        return out


@settings(max_examples=120, deadline=None)
@given(seconds=st.lists(st.integers(min_value=1, max_value=30), min_size=1, max_size=40), world=st.integers(min_value=1, max_value=4),
       rank_seed=st.integers(min_value=0, max_value=3), budget=st.integers(min_value=50, max_value=2000),
       tail_blocks=st.integers(min_value=1, max_value=6), tail_min=st.integers(min_value=1, max_value=600))
def test_contiguous_sharded_pushes_cover_the_share_exactly_once(seconds, world, rank_seed, budget, tail_blocks, tail_min):
    """Whatever the sub-batch budget and the tail blocking: the pushed blocks tile this rank's rows exactly once, in order, the
    last sub-batch (and only it) goes through the hidden-state path when it is split, and every row holds what its token
    produced."""
    import torch

    rank = rank_seed % world
    n_samples = [16000 * s + 37 for s in seconds]
    costs = [launcher.tokens_for_samples(n) for n in n_samples]
    offsets = np.concatenate([[0], np.cumsum(costs)])
    mine = launcher.contiguous_partition(costs, world)[rank]
    if not mine:
        return
    row_base, my_rows = int(offsets[mine[0]]), int(sum(costs[i] for i in mine))
    # the fake "audio": the first sample of every sub-batch slice must tell the encoder the global row it starts at, so fill each
    # utterance's samples with the global token offset of that utterance
    packed = torch.cat([torch.full((n_samples[i],), float(offsets[i])) for i in mine])
    g = _RecordingGather(int(offsets[-1]), 2, n_peers=world - 1)
    enc = _RowIdEncoder()
    emb, offs, got_mine = launcher.encode_contiguous_sharded(enc, packed, n_samples, rank, world, gather=g, tokens_per_call=budget,
                                                             out_dtype="float32", tail_blocks=tail_blocks, tail_min_rows=tail_min)
    assert got_mine == mine and list(offs) == list(offsets)
    pos = row_base
    for r0, n in g.pushed:  # contiguous, ordered, no gap, no overlap
        assert r0 == pos and n > 0
        pos += n
    assert pos == row_base + my_rows
    assert torch.equal(emb[row_base: row_base + my_rows, 0], torch.arange(row_base, row_base + my_rows, dtype=torch.float32))
    kinds = [k for k, _ in enc.calls]
    if "hidden" in kinds:  # only the LAST sub-batch is split, and only when there is a peer to push to
        assert world > 1 and kinds.index("hidden") == len([k for k in kinds if k == "full"]) and kinds.count("hidden") == 1
        blocks = [n for k, n in enc.calls if k == "block"]
        assert 2 <= len(blocks) <= tail_blocks and sum(blocks) == dict(enc.calls)["hidden"]


def test_bind_host_to_gpu_never_breaks_the_process():
    """Without NVML / a GPU the binding is a no-op that reports None; with one it returns a non-empty subset of the cores the
    process was allowed to use.  Either way the process keeps a usable affinity mask."""
    import os

    before = os.sched_getaffinity(0)
    cores = launcher.bind_host_to_gpu(0)
    after = os.sched_getaffinity(0)
    assert after and after <= before
    if cores is None:
        assert after == before
    else:
        assert set(cores) == after
    os.sched_setaffinity(0, before)
