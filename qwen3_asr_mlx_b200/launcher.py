"""Data-parallel launcher: shard utterance batches over the GPUs of one box.

One process per GPU (torchrun), weights replicated, NO collective on the forward path; the
only communication is the final gather of embeddings.  Utterances are
assigned to ranks by longest-processing-time-first over their token counts so that every
rank encodes about the same number of tokens.  The reference has no counterpart (it handles
one utterance at a time under a lock, model.py:145,239); results are defined as equal to a
per-utterance loop over the reference.
"""
from __future__ import annotations

import heapq
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

CHUNK_FRAMES = 100
TOKENS_PER_CHUNK = 13
HOP = 160


def bind_host_to_gpu(device: int) -> Optional[List[int]]:
    """Pin this process (and the pinned host buffers it allocates afterwards: first-touch NUMA placement) to the CPU cores
    NVML reports as local to ``device``.  With one process per GPU on a two-socket node, the per-step host<->device copies
    of the end-to-end path (123 MB in, 204 MB out per step and rank for config 2) otherwise cross the socket interconnect
    for half of the ranks.  Returns the core list, or None when NVML / the affinity call is unavailable (nothing changed).
    Honours CUDA_VISIBLE_DEVICES through the device's PCI bus id."""
    import os

    try:
        import pynvml

        pynvml.nvmlInit()
        bus = torch.cuda.get_device_properties(device).pci_bus_id if hasattr(torch.cuda.get_device_properties(device), "pci_bus_id") else None
        handle = None
        if bus is not None:
            for i in range(pynvml.nvmlDeviceGetCount()):
                hd = pynvml.nvmlDeviceGetHandleByIndex(i)
                if int(pynvml.nvmlDeviceGetPciInfo(hd).bus) == int(bus):
                    handle = hd
                    break
        if handle is None:
            handle = pynvml.nvmlDeviceGetHandleByIndex(device)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        allowed = os.sched_getaffinity(0)
        cores = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1 and 64 * w + b in allowed]
        if not cores:
            return None
        os.sched_setaffinity(0, cores)
        return cores
    except Exception:  # no NVML, container without the affinity syscall, ...: leave the placement to the OS
        return None


def conv_output_length(n: int) -> int:
    for _ in range(3):
        n = (n - 1) // 2 + 1
    return n


def tokens_for_samples(n_samples: int) -> int:
    """Audio tokens the encoder emits for an utterance of ``n_samples`` (encoder.py:258-293)."""
    frames = n_samples // HOP
    full, rem = divmod(frames, CHUNK_FRAMES)
    return full * TOKENS_PER_CHUNK + (conv_output_length(rem) if rem else 0)


def lpt_partition(costs: Sequence[int], world_size: int) -> List[List[int]]:
    """Greedy LPT: heaviest item first onto the currently lightest rank.  Deterministic
    (ties broken by index / rank), so every rank computes the same assignment locally."""
    order = sorted(range(len(costs)), key=lambda i: (-int(costs[i]), i))
    heap = [(0, r) for r in range(world_size)]
    heapq.heapify(heap)
    parts: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        load, r = heapq.heappop(heap)
        parts[r].append(i)
        heapq.heappush(heap, (load + int(costs[i]), r))
    for p in parts:
        p.sort()
    return parts


def contiguous_partition(costs: Sequence[int], world_size: int) -> List[List[int]]:
    """Token-balanced split of the utterance list into ``world_size`` CONTIGUOUS runs: rank r gets the utterances whose
    token prefix sum (mid-point) falls into the r-th equal share.  Imbalance is below one utterance per rank, like LPT on
    mixed-length data, but every rank's rows form ONE contiguous block of the gathered (original-order) matrix, so the final
    gather is a handful of large NVLink DMA copies (copy engines, no SM time) that overlap the next sub-batch's kernels
    instead of a row scatter.  Deterministic; every rank computes the same assignment locally."""
    c = np.asarray([int(v) for v in costs], dtype=np.int64)
    parts: List[List[int]] = [[] for _ in range(world_size)]
    total = int(c.sum())
    if total == 0:
        return parts
    mid = np.cumsum(c) - c / 2.0
    owner = np.minimum((mid * world_size / total).astype(np.int64), world_size - 1)
    owner = np.maximum.accumulate(owner)  # monotone even with zero-cost utterances
    for i, r in enumerate(owner):
        parts[int(r)].append(i)
    return parts


def split_by_budget(indices: Sequence[int], costs: Sequence[int], budget: int) -> List[List[int]]:
    """Split one rank's share into consecutive sub-batches of at most ``budget`` tokens each
    (bounds the activation workspace of one libqasr call)."""
    out: List[List[int]] = []
    cur: List[int] = []
    acc = 0
    for i in indices:
        c = int(costs[i])
        if cur and acc + c > budget:
            out.append(cur)
            cur, acc = [], 0
        cur.append(i)
        acc += c
    if cur:
        out.append(cur)
    return out


EncodeFn = Callable[[List[int]], Tuple[torch.Tensor, np.ndarray]]


def encode_sharded(encode_fn: EncodeFn, n_samples: Sequence[int], output_dim: int, rank: int, world_size: int,
                   tokens_per_call: int = 32768, gather: bool = True, group=None,
                   dtype: torch.dtype = torch.float32, peer_gather: Optional["PeerGather"] = None) -> Tuple[Optional[torch.Tensor], np.ndarray, List[int]]:
    """Encode utterances ``0..len(n_samples)-1`` data-parallel.

    ``encode_fn(indices)`` must return ``(embeddings (sum tokens, output_dim), token_offsets)`` for
    the utterances ``indices`` (in that order) on this rank's device.  Returns
    ``(all_embeddings or None, token_offsets (B+1,), my_indices)``: with ``gather=True`` every rank
    receives the embeddings of ALL utterances in original order (NCCL all-gather-v + one row gather, or, with a
    ``PeerGather``, one scatter kernel over NVLink peer memory).
    """
    costs = [tokens_for_samples(int(n)) for n in n_samples]
    parts = lpt_partition(costs, world_size)
    mine = parts[rank]
    pieces: List[torch.Tensor] = []
    for sub in split_by_budget(mine, costs, tokens_per_call):
        emb, toffs = encode_fn(sub)
        assert int(toffs[-1]) == sum(costs[i] for i in sub), "token count mismatch between host rule and library"
        pieces.append(emb)
    global_offsets = np.zeros(len(costs) + 1, dtype=np.int64)
    np.cumsum(np.asarray(costs, dtype=np.int64), out=global_offsets[1:])
    if not gather:
        local = torch.cat(pieces) if pieces else None
        return local, global_offsets, mine
    device = pieces[0].device if pieces else torch.device("cpu")
    local = torch.cat(pieces) if pieces else torch.zeros((0, output_dim), dtype=dtype, device=device)
    if peer_gather is not None:  # NVLink peer-memory scatter (transfer + order restore in one kernel) instead of NCCL
        return peer_gather.gather(local, parts, costs), global_offsets, mine
    return gather_embeddings(local, parts, costs, output_dim, rank, world_size, group), global_offsets, mine


def gather_start_rows(parts: List[List[int]], costs: np.ndarray, pad: int) -> np.ndarray:
    """Row of every utterance's first token inside the all-gathered buffer (rank r's share starts at ``r * pad`` and holds
    its utterances back to back in the order of ``parts[r]``)."""
    costs = np.asarray(costs, dtype=np.int64)
    start = np.zeros(len(costs), dtype=np.int64)
    for r, p in enumerate(parts):
        if p:
            idx = np.asarray(p, dtype=np.int64)
            pos = np.zeros(len(p), dtype=np.int64)
            np.cumsum(costs[idx][:-1], out=pos[1:])
            start[idx] = r * pad + pos
    return start


def gather_embeddings(local: torch.Tensor, parts: List[List[int]], costs: Sequence[int], output_dim: int, rank: int,
                      world_size: int, group=None) -> torch.Tensor:
    """Final gather (all-gather-v on a max-padded buffer) + restore of the original utterance order.

    The order is restored by ONE row gather (``index_select`` with a host-built row map) instead of a
    copy per utterance: at config 3 (4096 utterances) the per-utterance loop cost more than the collective."""
    per_rank = [sum(int(costs[i]) for i in p) for p in parts]
    total = sum(per_rank)
    costs_np = np.asarray([int(c) for c in costs], dtype=np.int64)
    offsets = np.zeros(len(costs) + 1, dtype=np.int64)
    np.cumsum(costs_np, out=offsets[1:])
    if world_size == 1:
        pad, flat = per_rank[0], local
    else:
        pad = max(per_rank)
        buf = local
        if local.shape[0] != pad:  # only the short ranks pad; the tail rows are never read back
            buf = torch.empty((pad, output_dim), dtype=local.dtype, device=local.device)
            buf[: local.shape[0]].copy_(local)
        flat = torch.empty((world_size * pad, output_dim), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(flat, buf.contiguous(), group=group)
    start = gather_start_rows(parts, costs_np, pad)
    if world_size == 1 and np.array_equal(start, offsets[:-1]):
        return flat  # already in original order
    src = np.repeat(start - offsets[:-1], costs_np) + np.arange(total, dtype=np.int64)
    return flat.index_select(0, torch.from_numpy(src).to(flat.device, non_blocking=True))


def scatter_dst_rows(my_indices: Sequence[int], costs: Sequence[int]) -> np.ndarray:
    """Final row (in the original-order gathered matrix) of every local row of a rank that holds the utterances
    ``my_indices`` back to back."""
    costs_np = np.asarray([int(c) for c in costs], dtype=np.int64)
    offsets = np.zeros(len(costs_np) + 1, dtype=np.int64)
    np.cumsum(costs_np, out=offsets[1:])
    if len(my_indices) == 0:
        return np.zeros(0, dtype=np.int64)
    idx = np.asarray(list(my_indices), dtype=np.int64)
    n = costs_np[idx]
    local_start = np.zeros(len(idx), dtype=np.int64)
    np.cumsum(n[:-1], out=local_start[1:])
    return np.repeat(offsets[idx] - local_start, n) + np.arange(int(n.sum()), dtype=np.int64)


class PeerGather:
    """Final gather through NVLink peer memory: every rank writes its rows into every rank's output buffer at their final
    positions with ONE kernel (``qasr_scatter_rows_to_peers``), so the transfer and the restore of the original utterance
    order are fused; no padded staging buffer, no NCCL collective on the data path (the only collective calls are the two
    device-side barriers of the symmetric-memory handle).  Buffers are torch symmetric memory (P2P-mapped by the rendezvous).

    One instance per process; ``capacity_rows`` bounds the gathered matrix.  Needs NVLink/P2P between the ranks' GPUs."""

    def __init__(self, capacity_rows: int, output_dim: int, dtype: torch.dtype = torch.bfloat16, group=None):
        import torch.distributed._symmetric_memory as symm

        from . import _lib

        self.lib = _lib.load()
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.dim, self.dtype = int(output_dim), dtype
        self.buf = symm.empty((int(capacity_rows), self.dim), dtype=dtype, device=torch.device("cuda", torch.cuda.current_device()))
        self.handle = symm.rendezvous(self.buf, self.group)
        self.row_bytes = self.dim * self.buf.element_size()
        if self.row_bytes % 16:
            raise ValueError("row size must be a multiple of 16 bytes")
        import ctypes

        self._ptrs = (ctypes.c_void_p * self.world)(*[int(p) for p in self.handle.buffer_ptrs])

    def gather(self, local: torch.Tensor, parts: List[List[int]], costs: Sequence[int]) -> torch.Tensor:
        """``local``: this rank's rows (its utterances ``parts[rank]`` back to back).  Returns the gathered matrix in original
        utterance order (a view of this rank's symmetric buffer, valid until the next ``gather``)."""
        import ctypes

        from . import _lib

        total = int(sum(int(c) for c in costs))
        if total > self.buf.shape[0]:
            raise ValueError(f"gathered matrix has {total} rows, capacity is {self.buf.shape[0]}")
        if local.dtype != self.dtype or local.shape[1] != self.dim:
            raise ValueError("local rows have the wrong dtype / width")
        local = local.contiguous()
        dst = torch.from_numpy(scatter_dst_rows(parts[self.rank], costs)).to(local.device, non_blocking=True)
        assert dst.numel() == local.shape[0], "token count mismatch between host rule and local rows"
        stream = torch.cuda.current_stream()
        self.handle.barrier(channel=0)  # every rank has finished reading the previous gather's result
        _lib.check(self.lib.qasr_scatter_rows_to_peers(ctypes.c_void_p(local.data_ptr()), local.shape[0], self.row_bytes,
                                                       ctypes.c_void_p(dst.data_ptr()), self._ptrs, self.world,
                                                       ctypes.c_void_p(stream.cuda_stream)))
        self.handle.barrier(channel=1)  # every rank's writes have landed
        return self.buf[:total]


class PeerBlockGather:
    """Final gather for CONTIGUOUS shares (``contiguous_partition`` / ``window_shares``): the gathered matrix lives in torch
    symmetric memory on every rank; a rank's kernels write its rows straight into its own copy at their final position
    (``rows()`` is passed as the encoder's ``out=``), and ``push()`` then copies that block into every peer's buffer with
    plain device-to-device copies over NVLink on side streams (copy engines: no SM is taken from the GEMMs), ordered after
    the producing kernels by an event.  The pushes of sub-batch k overlap the kernels of sub-batch k + 1; only the last
    block's transfer is exposed.  ``finish()`` joins the side streams and runs the symmetric-memory barrier, after which
    every rank holds all rows in original order.  No NCCL collective and no staging buffer on the data path."""

    def __init__(self, capacity_rows: int, output_dim: int, dtype: torch.dtype = torch.bfloat16, group=None, n_streams: int = 4):
        import torch.distributed._symmetric_memory as symm

        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.dim, self.dtype = int(output_dim), dtype
        dev = torch.device("cuda", torch.cuda.current_device())
        self.buf = symm.empty((int(capacity_rows), self.dim), dtype=dtype, device=dev)
        self.handle = symm.rendezvous(self.buf, self.group)
        # peers in a rotated order so that at any moment the ranks target different destinations
        self.peers = [(self.rank + k) % self.world for k in range(1, self.world)]
        self.peer_bufs = {p: self.handle.get_buffer(p, tuple(self.buf.shape), dtype) for p in self.peers}
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(max(1, min(n_streams, len(self.peers))))]
        self.bytes_pushed = 0
        self._pending = False
        self.trace = None  # set to [] to record (label, event) pairs: when each block was produced / had landed at the peers

    def begin(self) -> None:
        """Every rank has finished reading the previous result (its buffer may be overwritten)."""
        if self.trace is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record(torch.cuda.current_stream())
            self.trace.append(("begin", e))
        self.handle.barrier(channel=0)
        if self.trace is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record(torch.cuda.current_stream())
            self.trace.append(("begin barrier passed", e))

    def rows(self, row0: int, n: int) -> torch.Tensor:
        return self.buf[row0: row0 + n]

    def push(self, row0: int, n: int) -> None:
        """Copy rows [row0, row0 + n) of the local buffer (already queued on the current stream) to every peer."""
        if n <= 0 or not self.peers:
            return
        ev = torch.cuda.Event(enable_timing=self.trace is not None)
        ev.record(torch.cuda.current_stream())
        if self.trace is not None:
            self.trace.append((f"block@{row0} produced", ev))
        src = self.buf[row0: row0 + n]
        for k, p in enumerate(self.peers):
            st = self.streams[k % len(self.streams)]
            st.wait_event(ev)
            with torch.cuda.stream(st):
                self.peer_bufs[p][row0: row0 + n].copy_(src, non_blocking=True)
        if self.trace is not None:
            for i, st in enumerate(self.streams):
                e = torch.cuda.Event(enable_timing=True)
                e.record(st)
                self.trace.append((f"block@{row0} pushed (stream {i})", e))
        self.bytes_pushed += n * self.dim * self.buf.element_size() * len(self.peers)
        self._pending = True

    def finish(self, total_rows: int) -> torch.Tensor:
        cur = torch.cuda.current_stream()
        if self._pending:
            for st in self.streams:
                ev = torch.cuda.Event()
                ev.record(st)
                cur.wait_event(ev)
            self._pending = False
        if self.trace is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record(cur)
            self.trace.append(("local pushes joined", e))
        self.handle.barrier(channel=1)  # every rank's blocks have landed everywhere
        if self.trace is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record(cur)
            self.trace.append(("barrier passed", e))
        return self.buf[:total_rows]


def encode_contiguous_sharded(encoder, packed_audio: torch.Tensor, n_samples: Sequence[int], rank: int, world_size: int,
                              gather: Optional[PeerBlockGather] = None, tokens_per_call: int = 32768,
                              out_dtype: str = "bfloat16", tail_blocks: int = 4, tail_min_rows: int = 2048):
    """BASELINE config 3 on ``world_size`` GPUs with the gather hidden behind the compute.

    ``n_samples``: lengths of ALL utterances (original order); ``packed_audio``: THIS rank's share (``contiguous_partition``
    over token counts), packed back to back on the device.  The share is encoded in consecutive sub-batches of at most
    ``tokens_per_call`` tokens -- slices of ``packed_audio``, no host-side packing -- each writing its embeddings at their
    final rows of the symmetric buffer and pushed to the peers while the next sub-batch computes.
    The LAST sub-batch has nothing to hide its push behind, so its projector runs in ``tail_blocks`` row blocks
    (``encode_packed_audio_hidden`` + ``project_rows``: same bits) and every block is pushed as soon as it exists: only
    the last block's transfer (1 / ``tail_blocks`` of a sub-batch; blocks of at least ``tail_min_rows`` rows) stays exposed.
    Returns ``(embeddings, token_offsets (B+1,), my_indices)``; with ``gather=None`` the embeddings are this rank's rows only."""
    costs = [tokens_for_samples(int(n)) for n in n_samples]
    parts = contiguous_partition(costs, world_size)
    mine = parts[rank]
    offsets = np.zeros(len(costs) + 1, dtype=np.int64)
    np.cumsum(np.asarray(costs, dtype=np.int64), out=offsets[1:])
    total = int(offsets[-1])
    row_base = int(offsets[mine[0]]) if mine else 0
    my_rows = sum(costs[i] for i in mine)
    tdt = torch.bfloat16 if out_dtype in ("bfloat16", "bf16") else torch.float32
    if gather is not None:
        if total > gather.buf.shape[0] or gather.dtype != tdt:
            raise ValueError("gather buffer too small or of the wrong dtype")
        gather.begin()
        local = gather.rows(row_base, my_rows)
    else:
        local = torch.empty((my_rows, encoder.config.output_dim), dtype=tdt, device=packed_audio.device)
    sample_pos, row_pos = 0, 0
    # as many sub-batches as the budget requires, but of EQUAL size: no tiny tail call, and the last (exposed) push is 1/n
    n_sub = max(1, -(-my_rows // max(1, int(tokens_per_call))))
    even = -(-my_rows // n_sub) if my_rows else tokens_per_call
    longest = max((costs[i] for i in mine), default=0)
    subs = split_by_budget(mine, costs, min(int(tokens_per_call), even + longest))
    for k, sub in enumerate(subs):
        so = np.zeros(len(sub) + 1, dtype=np.int64)
        np.cumsum([int(n_samples[i]) for i in sub], out=so[1:])
        rows = sum(costs[i] for i in sub)
        audio = packed_audio[sample_pos: sample_pos + int(so[-1])]
        nblk = min(int(tail_blocks), rows // max(1, int(tail_min_rows))) if (gather is not None and gather.peers and k == len(subs) - 1) else 1
        if nblk > 1:
            n_hidden, _ = encoder.encode_packed_audio_hidden(audio, so)
            assert n_hidden == rows, "token count mismatch between host rule and library"
            step = -(-rows // nblk)
            step += (-step) % min(256, max(1, int(tail_min_rows)))  # whole 256-row GEMM tiles per block
            for r0 in range(0, rows, step):
                nr = min(step, rows - r0)
                encoder.project_rows(r0, local[row_pos + r0: row_pos + r0 + nr])
                gather.push(row_base + row_pos + r0, nr)
        else:
            _, toffs = encoder.encode_packed_audio(audio, so, out_dtype=out_dtype, out=local[row_pos: row_pos + rows])
            assert int(toffs[-1]) == rows, "token count mismatch between host rule and library"
            if gather is not None:
                gather.push(row_base + row_pos, rows)
        sample_pos += int(so[-1])
        row_pos += rows
    if gather is not None:
        return gather.finish(total), offsets, mine
    return local, offsets, mine


def window_shares(n_frames: int, world_size: int, window_frames: int = 800) -> List[Tuple[int, int]]:
    """Frame ranges [a, b) of ONE utterance for every rank, cut on the attention-window grid (n_window_infer = 800 frames =
    104 tokens, encoder.py:297-311) so that no window straddles two ranks; contiguous, as even as possible."""
    n_win = (n_frames + window_frames - 1) // window_frames
    base, extra = divmod(n_win, world_size)
    out, w = [], 0
    for r in range(world_size):
        k = base + (1 if r < extra else 0)
        out.append((min(w * window_frames, n_frames), min((w + k) * window_frames, n_frames)))
        w += k
    return out


def encode_long_sharded(encoder, mel: torch.Tensor, rank: int, world_size: int, group=None,
                        peer_gather=None, out_dtype: str = "float32") -> torch.Tensor:
    """One long utterance (BASELINE config 4, single pass: the reference's default ``chunk_duration`` does not split a
    20-minute file) on ``world_size`` GPUs.  ``mel`` is the utterance's full ``(128, T)`` log-mel -- every rank computes it
    from the waveform (it carries the utterance-wide max of audio.py:275; 0.4 ms for 20 minutes) -- and each rank encodes
    a contiguous share of whole attention windows: windows are independent given the mel, so the gathered result is
    bit-identical to a single-GPU pass.  Returns the ``(n_tokens, output_dim)`` embeddings on every rank."""
    T = int(mel.shape[1])
    shares = window_shares(T, world_size)
    a, b = shares[rank]
    cfg = encoder.config
    tdt = torch.bfloat16 if out_dtype in ("bfloat16", "bf16") else torch.float32
    costs = [tokens_for_samples((hi - lo) * HOP) for lo, hi in shares]  # a share is a run of full windows (+ the tail)
    if isinstance(peer_gather, PeerBlockGather):
        # shares are contiguous row blocks: encode, then DMA the block into every peer's buffer (copy engines over NVLink)
        row0 = sum(costs[:rank])
        peer_gather.begin()
        if b > a:
            emb, _ = encoder.encode_batch([mel[:, a:b].contiguous()], out_dtype=out_dtype)
            peer_gather.rows(row0, costs[rank]).copy_(emb.tensor)
            peer_gather.push(row0, costs[rank])
        return peer_gather.finish(sum(costs))
    if b > a:
        emb, _ = encoder.encode_batch([mel[:, a:b].contiguous()], out_dtype=out_dtype)
        local = emb.tensor
    else:
        local = torch.zeros((0, cfg.output_dim), dtype=tdt, device=mel.device)
    parts = [[r] for r in range(world_size)]
    if world_size == 1:
        return local
    if peer_gather is not None:
        return peer_gather.gather(local, parts, costs)
    return gather_embeddings(local, parts, costs, cfg.output_dim, rank, world_size, group)
