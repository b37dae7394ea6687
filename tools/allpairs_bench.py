#!/usr/bin/env python
"""BASELINE config 5: all-pairs MatchSiftData + RANSAC FindHomography over N_SETS keypoint sets of
8192 points, descriptor sets exchanged with an NCCL all-gather, pairs partitioned cyclically.

  python tools/allpairs_bench.py --sets 64                      (1 GPU)
  torchrun --nproc-per-node 8 ... tools/allpairs_bench.py --sets 256   (8 GPUs: the full config)

Each rank extracts its shard of the frames synth(1920,1080,3000+i) at peakThresh 0.5, keeps the first
8192 keypoints after a canonical sort, all-gathers the SiftPoint arrays (one exchange step), then
matches + fits its share of the unordered pairs.  Prints one JSON line (rank 0):
Mmatches/s = sum(n1) / t / 1e6, pairs/s, TFLOP/s = 2*128*sum(n1*n2)/t.
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sets", type=int, default=32)
    ap.add_argument("--points", type=int, default=8192)
    ap.add_argument("--loops", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=3, help="timed repetitions; the best is reported")
    ap.add_argument("--max-pairs", type=int, default=0, help="bound the number of pairs per rank (0 = all)")
    args = ap.parse_args()
    import torch
    import cusift_b200 as csb

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = csb.Context(local_rank, 2)
    prm = csb.make_params(5, 0.0, 0.5, 10.0, 0.0)
    P = args.points
    n_sets = args.sets
    per = (n_sets + world - 1) // world
    # rank-local extraction of its frames -> [per][P] SiftPoint records in a torch uint8 buffer
    local = torch.zeros((per, P, 588), dtype=torch.uint8, device="cuda")
    counts_local = torch.zeros(per, dtype=torch.int32, device="cuda")
    for k in range(per):
        f = rank * per + k
        if f >= n_sets:
            break
        pts = ctx.extract(csb.synth(1920, 1080, 3000 + f), prm, max_pts=32768)
        order = np.lexsort((pts["scale"], pts["coords2D"][:, 1], pts["coords2D"][:, 0], pts["subsampling"]))
        pts = np.ascontiguousarray(pts[order][:P])
        buf = torch.from_numpy(pts.view(np.uint8).reshape(len(pts), 588))
        local[k, : len(pts)].copy_(buf)
        counts_local[k] = len(pts)
    torch.cuda.synchronize()
    # the one exchange step: all-gather of the descriptor sets (and their sizes)
    if dist is not None:
        allsets = torch.empty((world * per, P, 588), dtype=torch.uint8, device="cuda")
        allcounts = torch.empty(world * per, dtype=torch.int32, device="cuda")
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.all_gather_into_tensor(allsets, local)
        dist.all_gather_into_tensor(allcounts, counts_local)
        e1.record()
        torch.cuda.synchronize()
        gather_ms = e0.elapsed_time(e1)
    else:
        allsets, allcounts, gather_ms = local, counts_local, 0.0
    counts = allcounts.cpu().numpy()[:n_sets]
    ptrs = [allsets[i].data_ptr() for i in range(n_sets)]
    pairs_all = csb.all_pairs(n_sets)
    mine = [(k, p) for k, p in enumerate(pairs_all) if k % world == rank]
    if args.max_pairs:
        mine = mine[: args.max_pairs]
    ids = [k for k, _ in mine]
    pairs = [p for _, p in mine]
    # warm-up on two pairs, then the timed region
    ctx.allpairs(ptrs, counts, pairs[:2], "l2", args.loops, 0.0, 0.80, 5.0, 1, ids[:2])
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    dt = 1e30
    for rep in range(args.reps):
        t0 = time.perf_counter()
        H, inl, nv = ctx.allpairs(ptrs, counts, pairs, "l2", args.loops, 0.0, 0.80, 5.0, 1, ids)
        torch.cuda.synchronize()
        t = time.perf_counter() - t0
        if os.environ.get("AP_VERBOSE"):
            print(f"rank {rank} rep {rep}: {t*1e3:.1f} ms for {len(pairs)} pairs", flush=True)
        dt = min(dt, t)
    tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
    work = torch.tensor([float(sum(counts[i] for i, _ in pairs)), float(sum(int(counts[i]) * int(counts[j]) for i, j in pairs)),
                         float(len(pairs)), float(inl.sum())], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
    if rank == 0:
        t = float(tt[0])
        q, qc, npairs, inl_sum = (float(x) for x in work)
        print(json.dumps({"metric": "all-pairs MatchSiftData + FindHomography", "n_gpus": world, "sets": n_sets,
                          "points_per_set": P, "pairs": int(npairs), "seconds": t, "Mmatches_per_s": q / t / 1e6,
                          "pairs_per_s": npairs / t, "useful_TFLOP_per_s": 2 * 128 * qc / t / 1e12,
                          "allgather_ms": gather_ms, "allgather_bytes_per_rank": int(per * P * 588),
                          "ransac_loops": args.loops, "mean_inliers": inl_sum / max(npairs, 1)}))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
