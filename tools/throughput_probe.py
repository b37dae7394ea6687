#!/usr/bin/env python
"""Throughput of csb_extract_batch on 1080p frames for several slot counts, with and without the
result download (development probe: where is the per-frame time going?)."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import cusift_b200 as csb  # noqa: E402

W, H, MAXPTS, NF, POOL = 1920, 1080, 16384, 384, 16
prm = csb.make_params(5, 0.0, 1.0, 10.0, 0.0)
imgs = [csb.synth(W, H, 1000 + i) for i in range(POOL)]
slot_list = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1, 2, 4, 8]
for slots in slot_list:
    ctx = csb.Context(0, slots)
    dev = [ctx.upload_image(im) for im in imgs]
    pitch = dev[0][1]
    dlist = [dev[k % POOL][0] for k in range(NF)]
    nbuf = max(2 * slots, 8)
    ds = [ctx.alloc(588 * MAXPTS) for _ in range(nbuf)]
    pins = [csb.PinnedArray(MAXPTS) for _ in range(nbuf)]
    dsl = [ds[k % nbuf] for k in range(NF)]
    for host_out in (True, False):
        hs = [pins[k % nbuf].ptr for k in range(NF)] if host_out else None
        for _ in range(2):
            ctx.extract_batch(dlist, W, H, pitch, prm, dsl, hs, MAXPTS)
        t0 = time.perf_counter()
        c = ctx.extract_batch(dlist, W, H, pitch, prm, dsl, hs, MAXPTS)
        dt = time.perf_counter() - t0
        print(f"slots {slots} host_out {host_out}: {NF/dt:8.1f} frames/s  {dt/NF*1e6:7.1f} us/frame  kp {c.mean():.0f}", flush=True)
    ctx.close()
