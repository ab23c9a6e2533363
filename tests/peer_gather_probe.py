"""Timing probe of launcher.PeerGather (not a pytest file; run under torchrun on >= 2 GPUs)."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qwen3_asr_mlx_b200 import launcher  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
costs = [launcher.tokens_for_samples(int(n)) for n in np.random.default_rng(20261018).integers(16000, 480001, size=4096)]
parts = launcher.lpt_partition(costs, world)
rows = sum(costs[i] for i in parts[rank])
localbuf = torch.randn(rows, 2048, device="cuda").bfloat16()
pg = launcher.PeerGather(sum(costs), 2048, dtype=torch.bfloat16)


def ev():
    return torch.cuda.Event(enable_timing=True)


for it in range(3):
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    e = [ev() for _ in range(2)]
    e[0].record()
    out = pg.gather(localbuf, parts, costs)
    e[1].record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    if rank == 0:
        print(f"iter {it}: device {e[0].elapsed_time(e[1]):.2f} ms, host wall {1e3 * (t1 - t0):.2f} ms, rows/rank {rows}", flush=True)
# NCCL path for comparison
for it in range(3):
    torch.cuda.synchronize(); dist.barrier()
    e = [ev() for _ in range(2)]
    e[0].record()
    ref = launcher.gather_embeddings(localbuf, parts, costs, 2048, rank, world)
    e[1].record()
    torch.cuda.synchronize()
    if rank == 0:
        print(f"nccl iter {it}: device {e[0].elapsed_time(e[1]):.2f} ms equal={bool(torch.equal(ref, out))}", flush=True)
dist.barrier()
dist.destroy_process_group()
