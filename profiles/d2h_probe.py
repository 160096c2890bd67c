#!/usr/bin/env python
"""Host-side ceiling of the end-to-end path: plain cudaMemcpyAsync device -> pinned host, all ranks at once.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/d2h_probe.py

Every rank copies a 256 MiB device buffer into its own pinned host buffer 20 times, all ranks started together; prints the
per-rank and the aggregate GB/s.  The e2e leg of bench.py moves 268 MB of frames per step per GPU over the same path, so
its frames/s cannot exceed aggregate / 4.19 MB."""
import os
import time

import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 256 << 20
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h.copy_(d, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
reps = 20
for _ in range(reps):
    h.copy_(d, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
gbs = reps * n / dt / 1e9
t = torch.tensor([gbs], dtype=torch.float64, device="cuda")
if world > 1:
    allv = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allv, t)
    vals = [float(v) for v in allv]
else:
    vals = [gbs]
if rank == 0:
    print("d2h pinned, %d rank(s) at once: per rank %s GB/s, aggregate %.1f GB/s -> e2e ceiling %.0f frames/s of 1024^2 RGBA" % (
        world, ", ".join("%.1f" % v for v in vals), sum(vals), sum(vals) * 1e9 / (1024 * 1024 * 4)))
if world > 1:
    dist.destroy_process_group()
