#!/usr/bin/env python
"""Runs N 1080p frames of the bench workload through the product library (ncu target)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import cusift_b200 as csb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
w, h = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1920, 1080)
thresh = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
ctx = csb.Context(0, 1)
img = csb.synth(w, h, 1000)
d_img, pitch = ctx.upload_image(img)
d_sift = ctx.alloc(588 * 131072)
pin = csb.PinnedArray(131072)
prm = csb.make_params(5, 0.0, thresh, 10.0, 0.0)
for _ in range(n):
    c = ctx.extract_batch([d_img], w, h, pitch, prm, [d_sift], [pin.ptr], 131072)
print("frames", n, "keypoints", int(c[0]))
ctx.close()
