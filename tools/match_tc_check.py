#!/usr/bin/env python
"""Development check of the tensor-core matcher against the CPU oracle (run under `timeout`)."""
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import cusift_b200 as csb  # noqa: E402
import parity_utils as PU  # noqa: E402
from oracle import oracle as O  # noqa: E402

ctx = csb.Context(0, 1)
L = csb.lib()
rng = np.random.default_rng(0)


def rand_set(n, seed):
    r = np.random.default_rng(seed)
    s = np.zeros(n, csb.SIFT_DTYPE)
    d = np.abs(r.standard_normal((n, 128))).astype(np.float32)
    s["data"] = d / np.linalg.norm(d, axis=1, keepdims=True)
    s["coords2D"] = r.uniform(0, 1000, (n, 2)).astype(np.float32)
    return s


def check(name, a, b, dist):
    t0 = time.time()
    ours = ctx.match(a, b, dist)
    t1 = time.time()
    orc = O.match(a, b, dist)
    bad = {f: int(np.sum(ours[f] != orc[f])) for f in ("score", "ambiguity", "match", "match_xpos", "match_ypos")}
    print(name, dist, len(a), len(b), "mismatches", bad, "redo_blocks", L.csb_match_redo_blocks(ctx.h), f"{t1-t0:.3f}s", flush=True)
    return sum(bad.values())


s1 = O.read_vlfeat_sift(PU.GOLDEN / "sift1.bin")
s2 = O.read_vlfeat_sift(PU.GOLDEN / "sift2.bin")
tot = 0
tot += check("fixture", s1, s2, "l2")
tot += check("fixture", s1, s2, "dot")
tot += check("rand", rand_set(256, 1), rand_set(256, 2), "l2")
tot += check("rand", rand_set(300, 3), rand_set(700, 4), "l2")
a, b = rand_set(1000, 5), rand_set(3000, 6)
b["data"][17] = b["data"][3]; b["data"][24] = b["data"][3]; a["data"][0] = b["data"][3]
b["data"][1000] = a["data"][7]; b["data"][2999] = a["data"][7]
tot += check("dups", a, b, "l2")
tot += check("dups", a, b, "dot")
g1, g2 = PU.golden_frames()
p = csb.make_params(6, 0.0, 0.1)
k1 = ctx.extract(PU.preblur(g1), p, max_pts=32768)
k2 = ctx.extract(PU.preblur(g2), p, max_pts=32768)
tot += check("c1", k1, k2, "l2")
big = rand_set(8192, 9)
t0 = time.time(); m = ctx.match(big, big, "l2"); print("8192 self", time.time() - t0, int((m["match"] != np.arange(8192)).sum()), L.csb_match_redo_blocks(ctx.h))
sub = O.match(big[:128], big, "l2")
tot += int(sum(np.sum(m[f][:128] != sub[f]) for f in ("score", "ambiguity", "match")))
# timing
d1, d2 = ctx.upload_sift(big), ctx.upload_sift(rand_set(8192, 10))
ctx.profile(True); ctx.profile_reset()
for _ in range(5):
    L.csb_match(ctx.h, d1, 8192, d2, 8192, 1, None)
print({k: round(v["total_ms"] / max(v["launches"], 1), 4) for k, v in ctx.profile_table().items() if k.startswith("match")})
print("TOTAL MISMATCHES", tot)
