#!/usr/bin/env python
"""Per-kernel device times (CUDA events, one 1080p frame in flight) + 8-slot throughput of the library
selected by CSB_LIB_PATH (kernel-variant experiments)."""
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import cusift_b200 as csb  # noqa: E402

W, H, MAXPTS = 1920, 1080, 16384
prm = csb.make_params(5, 0.0, 1.0, 10.0, 0.0)
POOL = 8
imgs = [csb.synth(W, H, 1000 + i) for i in range(POOL)]
ctx = csb.Context(0, 8)
dev = [ctx.upload_image(im) for im in imgs]
pitch = dev[0][1]
ds = [ctx.alloc(588 * MAXPTS) for _ in range(16)]
pins = [csb.PinnedArray(MAXPTS) for _ in range(16)]
for k in range(8):
    ctx.extract_batch([dev[k][0]], W, H, pitch, prm, [ds[0]], [pins[0].ptr], MAXPTS)
ctx.profile(True)
ctx.profile_reset()
for k in range(48):
    ctx.extract_batch([dev[k % POOL][0]], W, H, pitch, prm, [ds[0]], [pins[0].ptr], MAXPTS)
tab = ctx.profile_table()
ctx.profile(False)
grp = {"pyramid_o0": 0.0, "down_chain": 0.0, "pyramid_rest": 0.0, "blur_dog": 0.0, "find_points": 0.0, "orient_desc": 0.0, "copy_out": 0.0}
for name, v in tab.items():
    if v["launches"]:
        for g in grp:
            if name.startswith(g):
                grp[g] += v["total_ms"] * 1e3 / 48
o0 = {}
NF = 512
dl = [dev[k % POOL][0] for k in range(NF)]
dsl = [ds[k % 16] for k in range(NF)]
hs = [pins[k % 16].ptr for k in range(NF)]
best = 0.0
for _ in range(4):
    t0 = time.perf_counter()
    ctx.extract_batch(dl, W, H, pitch, prm, dsl, hs, MAXPTS)
    best = max(best, NF / (time.perf_counter() - t0))
print(os.environ.get("CSB_LIB_PATH", "default"), {k: round(v, 1) for k, v in grp.items()}, o0, f"throughput {best:.0f} fps", flush=True)
ctx.close()
