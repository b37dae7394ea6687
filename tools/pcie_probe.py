#!/usr/bin/env python
"""Aggregate host<->device copy bandwidth with all ranks copying at once (explains the multi-GPU e2e numbers).
  torchrun --nproc-per-node N tools/pcie_probe.py"""
import json
import os
import time

import torch
import torch.distributed as dist

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nbytes, reps = 64 << 20, 40
h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
out = {}
for name, (dst, src) in {"d2h": (h, d), "h2d": (d, h)}.items():
    for _ in range(3):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out[name] = {"per_gpu_GBs": nbytes * reps / float(t[0]) / 1e9, "aggregate_GBs": world * nbytes * reps / float(t[0]) / 1e9}
if rank == 0:
    print(json.dumps({"n_gpus": world, **out}))
if world > 1:
    dist.destroy_process_group()
