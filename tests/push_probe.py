"""Probe (not a pytest file): throughput of the NVLink DMA block pushes of launcher.PeerBlockGather, alone and while the
encoder is computing on the same GPU.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 tests/push_probe.py"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qwen3_asr_mlx_b200 import AudioEncoder, AudioEncoderConfig, launcher, weights  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = AudioEncoderConfig()
rows = 32768
for n_streams in (1, 2, 4, 7):
    g = launcher.PeerBlockGather(rows * 8, cfg.output_dim, dtype=torch.bfloat16, n_streams=n_streams)
    g.buf.fill_(1.0)
    torch.cuda.synchronize()
    dist.barrier()
    # emulate 7 peers on a 2-GPU box: push the same block 7 times to the one peer (different destination rows)
    peer = g.peers[0]

    def push_all(reps=7):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        src = g.buf[:rows]
        for k in range(reps):
            st = g.streams[k % len(g.streams)]
            st.wait_event(ev)
            with torch.cuda.stream(st):
                g.peer_bufs[peer][(k + 1) * rows: (k + 2) * rows].copy_(src, non_blocking=True)
        for st in g.streams:
            e = torch.cuda.Event()
            e.record(st)
            torch.cuda.current_stream().wait_event(e)

    push_all()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    push_all()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    gb = 7 * rows * cfg.output_dim * 2 / 1e9
    if rank == 0:
        print(f"streams={n_streams}: 7 x {rows} rows ({gb:.2f} GB) alone: {ms:.2f} ms = {gb / (ms / 1e3):.0f} GB/s", flush=True)
    del g
# while computing
enc = AudioEncoder(cfg, device=local)
enc.load_weights(weights.random_init(cfg, seed=1234))
from bench import make_workload  # noqa: E402
audio, soffs = make_workload(rank)
audio_dev = torch.from_numpy(audio).cuda()
out = torch.empty((64 * 390, cfg.output_dim), dtype=torch.bfloat16, device="cuda")
for _ in range(3):
    enc.encode_packed_audio(audio_dev, soffs, out_dtype="bfloat16", out=out)
torch.cuda.synchronize()
for n_streams in (2, 7):
    g = launcher.PeerBlockGather(rows * 8, cfg.output_dim, dtype=torch.bfloat16, n_streams=n_streams)
    g.buf.fill_(1.0)
    torch.cuda.synchronize()
    dist.barrier()
    peer = g.peers[0]
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    done = []
    e0.record()
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream())
    for k in range(7):
        st = g.streams[k % len(g.streams)]
        st.wait_event(ev)
        with torch.cuda.stream(st):
            g.peer_bufs[peer][(k + 1) * rows: (k + 2) * rows].copy_(g.buf[:rows], non_blocking=True)
    for st in g.streams:
        e = torch.cuda.Event(enable_timing=True)
        e.record(st)
        done.append(e)
    enc.encode_packed_audio(audio_dev, soffs, out_dtype="bfloat16", out=out)  # ~26 ms of kernels on the main stream
    e1.record()
    torch.cuda.synchronize()
    if rank == 0:
        print(f"streams={n_streams}: pushes done after {max(e0.elapsed_time(d) for d in done):.2f} ms while one encoder step ran in {e0.elapsed_time(e1):.2f} ms", flush=True)
    del g
dist.barrier()
dist.destroy_process_group()
