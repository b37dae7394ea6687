#!/usr/bin/env python
"""Soak test: many frames of varying sizes / parameters through one context, checking counts stay
reproducible and device memory does not creep (development aid)."""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import cusift_b200 as csb  # noqa: E402
import torch  # noqa: E402

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 20.0
ctx = csb.Context(0, 4)
cases = [(1920, 1080, 5, 1.0), (640, 480, 6, 0.1), (3840, 2160, 5, 2.0), (173, 131, 3, 0.5), (1280, 720, 4, 0.7)]
imgs = {c: csb.synth(c[0], c[1], 77 + i) for i, c in enumerate(cases)}
ref = {}
t0 = time.time()
n = 0
free0 = None
while time.time() - t0 < secs:
    c = cases[n % len(cases)]
    k = ctx.extract(imgs[c], csb.make_params(c[2], 0.0, c[3]), max_pts=65536)
    key = (len(k), float(np.sort(k["coords2D"][:, 0]).sum()))
    if c in ref:
        assert ref[c] == key, (c, ref[c], key)
    ref[c] = key
    n += 1
    if n % len(cases) == 0:                 # same workspace live at every snapshot
        free1 = torch.cuda.mem_get_info()[0]
        if n == 50:
            free0 = free1
print(f"soak ok: {n} frames in {secs:.0f} s, counts {[v[0] for v in ref.values()]}, free memory drift {(free0 - free1) / 1e6:.1f} MB")
ctx.close()
