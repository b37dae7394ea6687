#!/usr/bin/env python
"""Soak test of the matcher / all-pairs path: the same calls repeated for `secs` seconds, every result compared with
the first one (the tcgen05 scan, the rescoring kernel, the exact redo pass, the cluster ImproveHomography and the fused
RANSAC kernels must be deterministic and must not hang).  Development aid:  python tools/soak_match.py 30"""
import sys
import time
import zlib
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import cusift_b200 as csb  # noqa: E402

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 20.0


def rand_set(n, seed):
    r = np.random.default_rng(seed)
    s = np.zeros(n, csb.SIFT_DTYPE)
    d = np.abs(r.standard_normal((n, 128))).astype(np.float32)
    s["data"] = d / np.linalg.norm(d, axis=1, keepdims=True)
    s["coords2D"] = r.uniform(0, 1000, (n, 2)).astype(np.float32)
    return s


def digest(m):
    return zlib.crc32(np.ascontiguousarray(m["score"]).tobytes() + np.ascontiguousarray(m["match"]).tobytes()
                      + np.ascontiguousarray(m["ambiguity"]).tobytes())


ctx = csb.Context(0, 2)
a, b = rand_set(4096, 1), rand_set(5000, 2)
for c in range(40, 40 + 21 * 16, 16):
    b["data"][c] = a["data"][5]                      # one query whose list overflows: the redo pass runs every time
prm = csb.make_params(5, 0.0, 0.5)
sets = []
for k in range(4):
    pts = ctx.extract(csb.synth(640, 480, 3000 + k), prm, max_pts=8192)
    order = np.lexsort((pts["scale"], pts["coords2D"][:, 1], pts["coords2D"][:, 0], pts["subsampling"]))
    sets.append(np.ascontiguousarray(pts[order][:1536]))
dp = [ctx.upload_sift(s) for s in sets]
pairs = csb.all_pairs(len(sets))
ref_m = ref_ap = None
n = 0
t0 = time.time()
while time.time() - t0 < secs:
    m = ctx.match(a, b, "l2" if n % 2 == 0 else "dot")
    key = (n % 2, digest(m))
    if ref_m is None:
        ref_m = {}
    assert ref_m.setdefault(n % 2, key) == key, ("match differs", n)
    H, inl, nv, H2, nf = ctx.allpairs(dp, [len(s) for s in sets], pairs, "l2", 256, 0.0, 0.80, 5.0, 7, improve_loops=3)
    kap = (tuple(inl.tolist()), tuple(nv.tolist()), tuple(nf.tolist()), zlib.crc32(H.tobytes()), zlib.crc32(H2.tobytes()))
    if ref_ap is None:
        ref_ap = kap
    assert ref_ap == kap, ("all-pairs differs", n)
    n += 1
print(f"soak ok: {n} rounds in {secs:.0f} s (match 4096 x 5000 with a redo block, 6 pairs with RANSAC + ImproveHomography), "
      f"redo blocks {csb.lib().csb_match_redo_blocks(ctx.h)}")
ctx.close()
