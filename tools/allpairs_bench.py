#!/usr/bin/env python
"""BASELINE config 5: all-pairs MatchSiftData + RANSAC FindHomography (+ ImproveHomography) over N_SETS keypoint sets
of 8192 points through csb_allpairs_distributed: the sets are exchanged with NCCL all-gathers issued from the C++
library, the unordered pairs partitioned cyclically, the per-pair results all-gathered.

  python tools/allpairs_bench.py --sets 64                                          (1 GPU)
  torchrun --nproc-per-node 8 ... tools/allpairs_bench.py --sets 256                (8 GPUs: the full config)
  torchrun --nproc-per-node 2 ... tools/allpairs_bench.py --sets 8 --points 2048 --check   (results == 1-rank run)

Each rank extracts its shard of the frames synth(1920,1080,3000+i) at peakThresh 0.5 and keeps the first `points`
keypoints after a canonical sort.  Prints one JSON line (rank 0): Mmatches/s = sum(n1) / t / 1e6, pairs/s,
TFLOP/s = 2*128*sum(n1*n2)/t, where t is the max over ranks of the whole csb_allpairs_distributed call.
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
    ap.add_argument("--improve", type=int, default=0, help="ImproveHomography loops appended to every pair")
    ap.add_argument("--reps", type=int, default=2, help="timed repetitions; the best is reported")
    ap.add_argument("--size", default="1920x1080")
    ap.add_argument("--check", action="store_true", help="rank 0 recomputes every pair alone and compares")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import cusift_b200 as csb

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("gloo")          # plumbing only (id broadcast, barriers, timing reduction): the data path is NCCL from C++
    W, H = (int(x) for x in args.size.split("x"))
    ctx = csb.Context(local_rank, 2)
    prm = csb.make_params(5, 0.0, 0.5, 10.0, 0.0)
    P, n_sets = args.points, args.sets
    assert n_sets % world == 0, "sets must divide evenly over the ranks"
    per = n_sets // world
    sets, dptrs, counts = [], [], []
    for k in range(per):
        pts = ctx.extract(csb.synth(W, H, 3000 + rank * per + k), prm, max_pts=32768)
        order = np.lexsort((pts["scale"], pts["coords2D"][:, 1], pts["coords2D"][:, 0], pts["subsampling"]))
        pts = np.ascontiguousarray(pts[order][:P])
        sets.append(pts)
        dptrs.append(ctx.upload_sift(pts))
        counts.append(len(pts))

    def bcast(raw):
        if world == 1:
            return raw
        obj = [raw]
        dist.broadcast_object_list(obj, src=0)
        return obj[0]

    comm = ctx.nccl_comm(rank, world, bcast) if world > 1 else None
    kw = dict(distance="l2", num_loops=args.loops, min_score=0.0, max_ambiguity=0.80, thresh=5.0, seed=1,
              improve_loops=args.improve, improve_thresh=3.0)
    res = ctx.allpairs_distributed(comm, rank, world, dptrs, counts, P, **kw)      # warm-up (allocations, NCCL channels)
    best, tm_best = 1e30, None
    for _ in range(args.reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = ctx.allpairs_distributed(comm, rank, world, dptrs, counts, P, **kw)
        t = time.perf_counter() - t0
        tt = torch.tensor([t], dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        if float(tt[0]) < best:
            best, tm_best = float(tt[0]), res["timings_ms"].copy()
    all_counts = [None] * world
    if world > 1:
        dist.all_gather_object(all_counts, counts)
    else:
        all_counts = [counts]
    cnt = np.array([c for r in all_counts for c in r], np.int64)
    pairs = csb.all_pairs(n_sets)
    ok = None
    if args.check:
        # every rank must hold identical, complete results; rank 0 additionally recomputes all pairs on its own
        h = [None] * world
        digest = (res["H"].tobytes(), res["inliers"].tobytes(), res["n_valid"].tobytes())
        if world > 1:
            dist.all_gather_object(h, digest)
            gathered = [None] * world
            dist.all_gather_object(gathered, sets)
        else:
            h, gathered = [digest], [sets]
        ok = all(x == h[0] for x in h)
        if rank == 0:
            allsets = [s for r in gathered for s in r]
            d_all = [ctx.upload_sift(s) for s in allsets]
            solo = ctx.allpairs(d_all, [len(s) for s in allsets], pairs, "l2", args.loops, 0.0, 0.80, 5.0, 1, None,
                                improve_loops=args.improve, improve_thresh=3.0)
            ok = ok and np.array_equal(solo[1], res["inliers"]) and np.array_equal(solo[2], res["n_valid"]) and \
                np.array_equal(solo[0], res["H"])
            if args.improve:
                ok = ok and np.array_equal(solo[4], res["num_fit"]) and np.allclose(solo[3], res["H_improved"], rtol=1e-6, atol=1e-6)
            for d in d_all:
                ctx.free(d)
    if rank == 0:
        q = float(sum(cnt[i] for i, _ in pairs))
        qc = float(sum(int(cnt[i]) * int(cnt[j]) for i, j in pairs))
        line = {"metric": "all-pairs MatchSiftData + FindHomography", "n_gpus": world, "sets": n_sets, "points_per_set": P,
                "pairs": len(pairs), "seconds": best, "Mmatches_per_s": q / best / 1e6, "pairs_per_s": len(pairs) / best,
                "useful_TFLOP_per_s": 2 * 128 * qc / best / 1e12, "ransac_loops": args.loops, "improve_loops": args.improve,
                "exchange": "ncclAllGather of SiftPoint arrays from C++ (csb_allpairs_distributed), one per local set index, "
                            "underneath the rank-local pairs" if world > 1 else "none (1 GPU)",
                "exchange_bytes_per_rank": int(per * P * 588), "rank0_timings_ms": {
                    "local_pairs_and_queueing": float(tm_best[0]), "wait_for_exchange": float(tm_best[1]),
                    "remaining_pairs": float(tm_best[2]), "result_exchange": float(tm_best[3])},
                "mean_inliers": float(res["inliers"].mean()), "check_vs_single_rank": ok}
        print(json.dumps(line))
        if args.out:
            Path(args.out).write_text(json.dumps(line) + "\n")
    if comm is not None:
        ctx.nccl_comm_destroy(comm)
    for d in dptrs:
        ctx.free(d)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
