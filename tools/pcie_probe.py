#!/usr/bin/env python
"""Host → device bandwidth of N GPUs at once, OUTSIDE the library: plain cudaMemcpyAsync from pinned memory (torch
copy_ with non_blocking), one process per GPU (torchrun), all ranks copying simultaneously.  Answers whether the
end-to-end leg of bench.py (which uploads every rank's shard from pinned host memory) is limited by the library or by
the box's host path (PCIe switches / root complexes / host memory): per-GPU GB/s alone vs per-GPU GB/s with all N busy.

    python -m torch.distributed.run --nproc-per-node N tools/pcie_probe.py
"""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30  # 1 GiB per copy
host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
host.fill_(1)
devb = torch.empty(n, dtype=torch.uint8, device="cuda")


def h2d(reps=8):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        devb.copy_(host, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return reps * n / (e0.elapsed_time(e1) * 1e-3) / 1e9


def d2h(reps=8):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        host.copy_(devb, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return reps * n / (e0.elapsed_time(e1) * 1e-3) / 1e9


h2d(2)
res = {}
# one GPU at a time
for r in range(world):
    if world > 1:
        dist.barrier()
    if r == rank:
        res["alone_h2d"], res["alone_d2h"] = h2d(), d2h()
if world > 1:
    dist.barrier()
# subsets busy at once: the first k ranks copy simultaneously
k = 2
while k <= world:
    dist.barrier()
    if rank < k:
        res[f"h2d_with_{k}_busy"] = h2d()
    dist.barrier()
    if rank < k:
        res[f"d2h_with_{k}_busy"] = d2h()
    k *= 2
line = f"rank {rank} (cuda:{local}): " + "  ".join(f"{a}={b:.1f}" for a, b in res.items())
if world > 1:
    out = [None] * world
    dist.all_gather_object(out, line)
    if rank == 0:
        print("\n".join(out))
        tot = [None] * world
    dist.barrier()
    dist.destroy_process_group()
else:
    print(line)
