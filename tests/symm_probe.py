"""Probe (not a pytest file): does torch symmetric memory give P2P-mapped peer buffers on this box?
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tests/symm_probe.py"""
import os
import time

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t = symm.empty((200000, 2048), dtype=torch.bfloat16, device=f"cuda:{local}")
hdl = symm.rendezvous(t, dist.group.WORLD)
print(rank, "rendezvous ok", type(hdl).__name__, [hex(p) for p in hdl.buffer_ptrs][:4], flush=True)
t.fill_(rank)
hdl.barrier()
peer = hdl.get_buffer((rank + 1) % world, t.shape, t.dtype)
peer[:10] = 100 + rank
hdl.barrier()
torch.cuda.synchronize()
print(rank, "first rows", float(t[0, 0]), float(t[20, 0]), flush=True)
src = torch.randn(100000, 2048, device="cuda").bfloat16()
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    peer[100000:200000].copy_(src)
e1.record()
torch.cuda.synchronize()
print(rank, "peer copy GB/s", 5 * src.numel() * 2 / (e0.elapsed_time(e1) * 1e-3) / 1e9, flush=True)
dist.barrier()
dist.destroy_process_group()
