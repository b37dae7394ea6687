"""CPU: multi-GPU host logic — frame / pair sharding, and a world_size-2 gloo run
of the same rank-partition + gather plumbing bench.py uses (no data-path collective)."""
import os
import subprocess
import sys
import textwrap
from pathlib import Path

import pytest

from cusift_b200 import all_pairs, pair_index, shard_frames, shard_pairs

ROOT = Path(__file__).resolve().parent.parent


def test_shard_frames_partition():
    for n, world in ((4096, 8), (10, 3), (5, 8), (0, 2)):
        shards = [shard_frames(n, r, world) for r in range(world)]
        flat = sorted(f for s in shards for f in s)
        assert flat == list(range(n))
        assert max(map(len, shards)) - min(map(len, shards)) <= 1
    with pytest.raises(ValueError):
        shard_frames(4, 4, 4)


def test_shard_pairs_partition():
    n = 16
    pairs = all_pairs(n)
    assert len(pairs) == n * (n - 1) // 2
    assert [pair_index(i, j, n) for i, j in pairs] == list(range(len(pairs)))
    for world in (1, 2, 8):
        shards = [shard_pairs(n, r, world) for r in range(world)]
        assert sorted(p for s in shards for p in s) == pairs
        assert max(map(len, shards)) - min(map(len, shards)) <= 1
    assert len(all_pairs(256)) == 32640                     # BASELINE config 5


WORKER = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, {root!r})
    import torch, torch.distributed as dist
    from cusift_b200 import shard_frames
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
    rank = dist.get_rank()
    mine = shard_frames(37, rank, 2)
    # stand-in for per-frame keypoint counts produced by the rank-local extraction
    counts = torch.tensor([1000 + f for f in mine] + [0] * (19 - len(mine)), dtype=torch.int64)
    gathered = [torch.zeros(19, dtype=torch.int64) for _ in range(2)]
    dist.all_gather(gathered, counts)
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)                # max-over-ranks timing reduction
    if rank == 0:
        total = int(sum(int(g.sum()) for g in gathered))
        print(json.dumps({{"total": total, "tmax": float(t)}}))
    dist.barrier()
    dist.destroy_process_group()
""")


def test_gloo_world2_gather(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=str(ROOT), port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                              text=True) for r in range(2)]
    outs = [p.communicate(timeout=240) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    import json
    res = json.loads(outs[0][0].strip().splitlines()[-1])
    assert res["total"] == sum(1000 + f for f in range(37))
    assert res["tmax"] == 2.0
