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
# both directions at once in the ratio of the e2e arm (8.3 MB frame up, ~4.5 MB of SiftPoints down per frame)
up_b, dn_b = 64 << 20, 35 << 20
hu = torch.empty(up_b, dtype=torch.uint8).pin_memory()
du = torch.empty(up_b, dtype=torch.uint8, device="cuda")
hd = torch.empty(dn_b, dtype=torch.uint8).pin_memory()
dd = torch.empty(dn_b, dtype=torch.uint8, device="cuda")
s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
for it in range(2):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        with torch.cuda.stream(s_up):
            du.copy_(hu, non_blocking=True)
        with torch.cuda.stream(s_dn):
            hd.copy_(dd, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
t = torch.tensor([dt], device="cuda", dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
out["bidirectional_8.3up_4.5down"] = {"h2d_aggregate_GBs": world * up_b * reps / float(t[0]) / 1e9,
                                      "d2h_aggregate_GBs": world * dn_b * reps / float(t[0]) / 1e9,
                                      "frames_per_s_ceiling_fp32_aos": world * reps * (up_b / (1920 * 1080 * 4)) / float(t[0])}
if rank == 0:
    print(json.dumps({"n_gpus": world, **out}))
if world > 1:
    dist.destroy_process_group()
