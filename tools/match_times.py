#!/usr/bin/env python
"""Per-kernel CUDA-event times of one 8192 x 8192 csb_match (kernel-variant experiments; CSB_LIB_PATH selects the library)."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import cusift_b200 as csb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192


def rand_set(n, seed):
    r = np.random.default_rng(seed)
    s = np.zeros(n, csb.SIFT_DTYPE)
    d = np.abs(r.standard_normal((n, 128))).astype(np.float32)
    s["data"] = d / np.linalg.norm(d, axis=1, keepdims=True)
    return s


ctx = csb.Context(0, 1)
L = csb.lib()
d1, d2 = ctx.upload_sift(rand_set(n, 1)), ctx.upload_sift(rand_set(n, 2))
for _ in range(3):
    L.csb_match(ctx.h, d1, n, d2, n, 1, None)
ctx.profile(True)
ctx.profile_reset()
reps = 20
for _ in range(reps):
    L.csb_match(ctx.h, d1, n, d2, n, 1, None)
t = ctx.profile_table()
print({k: round(v["total_ms"] * 1e3 / reps, 2) for k, v in t.items() if k.startswith("match")}, "redo blocks", L.csb_match_redo_blocks(ctx.h))
ctx.close()
