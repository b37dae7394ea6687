#!/usr/bin/env python
"""How many 16-query blocks does the tensor-core matcher hand to the exact kernel on extracted (synthetic-frame) keypoint sets?"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import cusift_b200 as csb  # noqa: E402

ctx = csb.Context(0, 2)
prm = csb.make_params(5, 0.0, 0.5, 10.0, 0.0)
P = 8192
sets = []
for k in range(3):
    pts = ctx.extract(csb.synth(1920, 1080, 3000 + k), prm, max_pts=32768)
    order = np.lexsort((pts["scale"], pts["coords2D"][:, 1], pts["coords2D"][:, 0], pts["subsampling"]))
    sets.append(np.ascontiguousarray(pts[order][:P]))
L = csb.lib()
d = [ctx.upload_sift(s) for s in sets]
for i, j in ((0, 1), (0, 2), (1, 2)):
    b = L.csb_match_redo_blocks(ctx.h)
    L.csb_match(ctx.h, d[i], len(sets[i]), d[j], len(sets[j]), 1, None)
    print("pair", i, j, "n", len(sets[i]), len(sets[j]), "redo blocks", L.csb_match_redo_blocks(ctx.h) - b, "of", (len(sets[i]) + 15) // 16)
ctx.close()
